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


def test_named_workloads_are_deterministic_and_area_exact():
    """bench_workloads.make_workload: the five BASELINE configs — fixed slice lists (seed 42), closed-form areas equal to the
    dense mask areas on reduced sizes, and the C2 sliding-window decomposition equal to the direct definition."""
    import torch

    from bench_workloads import make_workload, sample_doc_lengths
    from oracle import ffa_oracle as orc

    assert make_workload("C1")["area"] == 4096 * 4097 // 2 and make_workload("C1")["fwd_only"]
    assert make_workload("C3")["area"] == 2147516416 and make_workload("C4")["area"] == 34359869440
    c5 = make_workload("C5")
    assert c5["num_slices"] == 1854 and c5["num_heads_q"] == 64 and c5["num_heads_kv"] == 8 and c5["seq_len"] == 1 << 20
    assert make_workload("C5")["q_ranges"] == c5["q_ranges"]  # deterministic
    for cfg in ("C2", "C5"):
        w = make_workload(cfg, seqlen=8192)
        T = w["seq_len"]
        m = orc.mask_from_slices(w["q_ranges"], w["k_ranges"], w["types"], T, T)
        assert int(m.sum()) == w["area"], cfg
    w = make_workload("C2", seqlen=8192)
    T, lens = w["seq_len"], sample_doc_lengths(8192, 8192 // 4)
    ref = torch.zeros(T, T, dtype=torch.bool)
    i, j, off = torch.arange(T)[:, None], torch.arange(T)[None, :], 0
    for d, n in enumerate(lens):
        blk = (i >= off) & (i < off + n) & (j >= off) & (j <= i)
        if d % 2 == 1:
            blk &= j >= i - 1024
        ref |= blk
        off += n
    assert torch.equal(orc.mask_from_slices(w["q_ranges"], w["k_ranges"], w["types"], T, T), ref)
