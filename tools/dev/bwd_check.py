"""Developer harness: backward kernel vs the oracle on a ladder of cases, then a first timing."""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def run_case(name, T, TK, HQ, HK, qr, kr, ty, dtype, dkv_f32=False, seed=0):
    import torch
    from magiattention_b200 import _lib
    from oracle import ffa_oracle as orc

    torch.manual_seed(seed)
    dev = torch.device("cuda:0")
    D = 128
    print(f"      ({name}: start)", flush=True)
    q = torch.randn(T, HQ, D, dtype=dtype, device=dev)
    k = torch.randn(TK, HK, D, dtype=dtype, device=dev)
    v = torch.randn(TK, HK, D, dtype=dtype, device=dev)
    do = torch.randn(T, HQ, D, dtype=dtype, device=dev)
    qr_t = torch.tensor(qr, dtype=torch.int32, device=dev).reshape(-1, 2)
    kr_t = torch.tensor(kr, dtype=torch.int32, device=dev).reshape(-1, 2)
    ty_t = torch.tensor(ty, dtype=torch.int32, device=dev)
    out = torch.empty(T, HQ, D, dtype=dtype, device=dev)
    lse = torch.empty(T, HQ, dtype=torch.float32, device=dev)
    scale = D ** -0.5
    _lib.ffa_fwd(q, k, v, out, lse, qr_t, kr_t, ty_t, scale)
    torch.cuda.synchronize()
    print(f"      ({name}: fwd done)", flush=True)
    dq = torch.zeros(T, HQ, D, dtype=torch.float32, device=dev)
    dkv_dt = torch.float32 if dkv_f32 else dtype
    dk = torch.full((TK, HK, D), float("nan"), dtype=dkv_dt, device=dev)
    dv = torch.full((TK, HK, D), float("nan"), dtype=dkv_dt, device=dev)
    _lib.ffa_bwd(do, q, k, v, out, lse, dq, dk, dv, qr_t, kr_t, ty_t, scale)
    torch.cuda.synchronize()
    print(f"      ({name}: kernels done, checking)", flush=True)
    mask = orc.mask_from_slices(qr, kr, ty, T, TK)
    qc, kc, vc, doc = q, k, v, do  # oracle math on the GPU (fp64 / bf16 torch ops): same code, faster
    ro, rl = orc.attn_fwd(qc, kc, vc, mask, scale, high_precision=True)
    ref = orc.attn_bwd(doc, qc, kc, vc, ro.double(), rl.double(), mask, scale, high_precision=True)
    lo, ll = orc.attn_fwd(qc, kc, vc, mask, scale, high_precision=False)
    lp = orc.attn_bwd(doc, qc, kc, vc, lo, ll, mask, scale, high_precision=False)
    ok = True
    msgs = []
    for nm, got, r, l in zip(("dq", "dk", "dv"), (dq, dk, dv), ref, lp):
        g = got.double()
        err = (g - r.double()).abs().max().item()
        err_lp = (l.double() - r.double()).abs().max().item()
        nan = bool(torch.isnan(g).any())
        good = (not nan) and err <= 2 * err_lp + 1e-5
        ok &= good
        msgs.append(f"{nm}: {err:.3e} (lp {err_lp:.3e}){'' if good else ' <-- BAD'}{' NaN' if nan else ''}")
        if not good:
            bad = ((g - r.double()).abs().amax(dim=(1, 2)) > 2 * err_lp + 1e-5).nonzero().flatten()
            msgs.append(f"[bad rows {bad[:8].tolist()} n={bad.numel()}]")
    print(f"[{'PASS' if ok else 'FAIL'}] {name:34s} T={T} TK={TK} h={HQ}/{HK}: " + " ".join(msgs), flush=True)
    return ok


def stage_A():
    import torch
    bf = torch.bfloat16
    ok = True
    ok &= run_case("full 128x128", 128, 128, 1, 1, [[0, 128]], [[0, 128]], [0], bf)
    ok &= run_case("full 256x128 (2 q blk)", 256, 128, 1, 1, [[0, 256]], [[0, 128]], [0], bf)
    ok &= run_case("full 128x256 (2 kv tiles)", 128, 256, 1, 1, [[0, 128]], [[0, 256]], [0], bf)
    ok &= run_case("full 640x384 GQA 4/2", 640, 384, 4, 2, [[0, 640]], [[0, 384]], [0], bf)
    ok &= run_case("full ragged 300x200", 300, 200, 2, 1, [[0, 300]], [[0, 200]], [0], bf)
    ok &= run_case("causal 1024", 1024, 1024, 2, 1, [[0, 1024]], [[0, 1024]], [1], bf)
    ok &= run_case("causal 1024 fp16 f32dkv", 1024, 1024, 2, 2, [[0, 1024]], [[0, 1024]], [1], torch.float16, dkv_f32=True)
    return ok


def stage_B():
    import torch
    from mask_configs import mask_configs
    ok = True
    for cfg in mask_configs(1):
        T = cfg["seqlen"]
        ok &= run_case(cfg["name"], T, T, 2, 1, cfg["q_ranges"], cfg["k_ranges"], cfg["types"], torch.bfloat16)
    return ok


def stage_C():
    import torch
    from magiattention_b200 import _lib
    dev = torch.device("cuda:0")
    D = int(os.environ.get("FFA_CHECK_D", "128"))
    for name, T, H, ty in (("causal 4K h16", 4096, 16, 1), ("causal 16K h16", 16384, 16, 1),
                           ("full 8K h16", 8192, 16, 0), ("causal 32K h8", 32768, 8, 1)):
        q, k, v, do = (torch.randn(T, H, D, dtype=torch.bfloat16, device=dev) for _ in range(4))
        out = torch.empty_like(q)
        lse = torch.empty(T, H, dtype=torch.float32, device=dev)
        qr = torch.tensor([[0, T]], dtype=torch.int32, device=dev)
        ty_t = torch.tensor([ty], dtype=torch.int32, device=dev)
        _lib.ffa_fwd(q, k, v, out, lse, qr, qr, ty_t, D ** -0.5)
        dq = torch.zeros(T, H, D, dtype=torch.float32, device=dev)
        dk, dv = torch.empty_like(k), torch.empty_like(v)
        for _ in range(2):
            _lib.ffa_bwd(do, q, k, v, out, lse, dq, dk, dv, qr, qr, ty_t, D ** -0.5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3
        e0.record()
        for _ in range(n):
            _lib.ffa_bwd(do, q, k, v, out, lse, dq, dk, dv, qr, qr, ty_t, D ** -0.5)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        area = T * T if ty == 0 else T * (T + 1) // 2
        fl = 2.5 * 4.0 * area * H * D
        print(f"[PERF] bwd {name:20s}: {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s", flush=True)
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", default=None)
    a = ap.parse_args()
    if a.stage:
        ok = {"A": stage_A, "B": stage_B, "C": stage_C}[a.stage]()
        sys.exit(0 if ok else 1)
    rc = 0
    for st in "ABC":
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", st], timeout=150)
            print(f"== stage {st}: rc={r.returncode} ({time.time() - t0:.0f}s)", flush=True)
            rc |= r.returncode
        except subprocess.TimeoutExpired:
            print(f"== stage {st}: TIMEOUT", flush=True)
            rc |= 1
    sys.exit(rc)
