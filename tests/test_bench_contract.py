"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the agreed keys, and the
file never imports the oracle outside the cpu_baseline / reference legs."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--gpus", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.strip().splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "TFLOP/s" and d["higher_is_better"] is True
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["value"] > 0


def test_reference_arm_nonzero_rank_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--gpus", "2"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_slices_area_matches_oracle_mask_area():
    """bench.py counts algorithmic FLOPs from closed-form slice areas; pin them to the oracle's dense masks."""
    import bench
    from mask_configs import mask_configs
    from oracle import ffa_oracle as orc

    for c in mask_configs(1):
        assert bench.slices_area(c["q_ranges"], c["k_ranges"], c["types"]) == orc.mask_area(c["q_ranges"], c["k_ranges"], c["types"])
    assert bench.slices_area([[0, 4096]], [[0, 4096]], [1]) == bench.causal_area(4096)
