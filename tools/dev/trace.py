"""Print the in-kernel timelines recorded by a -DFFA_TRACE build (B200_FFA_LIB=.../trace.so): clocks relative to the first
MMA-warp stamp of each iteration, mean over the recorded iterations, then a few raw iterations."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from magiattention_b200 import _lib  # noqa: E402

FWD = {0: ["pv0_begin", "pv0_Plo_ok", "pv0_Phi_ok", "s0_issued", "pv1_begin", "pv1_Plo_ok", "pv1_Phi_ok", "s1_issued"],
       1: ["sm0_wait_S", "sm0_S_ok", "sm0_ld_done", "sm0_Plo", "sm0_Phi"],
       2: ["sm1_wait_S", "sm1_S_ok", "sm1_ld_done", "sm1_Plo", "sm1_Phi"],
       3: ["tma_K_wait", "tma_K_go", "tma_V_wait", "tma_V_go"]}
BWD = {0: ["mma_wait_P", "mma_P_ok(dV,S+1)", "mma_wait_DS", "mma_DS_ok(dK,dQ)", "mma_wait_drain", "mma_drain_ok", "mma_dO_ok(dP+1)"],
       1: ["c_wait_S", "c_S_ok", "c_P_arrive", "c_DP_ok", "c_DQprev_ok", "c_DS_arrive"],
       2: ["d_wait_DQ", "d_DQ_ok", "d_drained", "d_end"],
       3: ["t_wait_Qempty", "t_Q_go", "t_dO_go"]}


def dump(name, table, fn):
    buf = np.zeros((4, 64, 8), dtype=np.uint64)
    assert fn(buf.ctypes.data_as(C.c_void_p)) == 0
    t = buf.astype(np.int64)
    base = t[0, :, 0:1]  # first MMA-warp stamp of each iteration
    valid = (t[0, :, 0] > 0)
    per_it = np.diff(t[0, valid, 0])
    print(f"== {name}: iteration period (MMA-warp stamp 0): mean {per_it.mean():.0f} clk, min {per_it.min()}, max {per_it.max()}")
    for role, names in table.items():
        rel = (t[role, :, :len(names)] - base)[valid]
        ok = (t[role, :, :len(names)][valid] > 0).all(axis=1)
        m = rel[ok].mean(axis=0)
        print(f"  role {role}: " + "  ".join(f"{n}={v:+.0f}" for n, v in zip(names, m)))
    for it in (10, 11, 12):
        print(f"  raw it {it}: " + " | ".join(" ".join(str(int(t[r, it, e] - t[0, it, 0])) for e in range(len(table[r]))) for r in table))


def main():
    dev = torch.device("cuda:0")
    T, H, D = 32768, 1, 128
    q, k, v, do = (torch.randn(T, H, D, dtype=torch.bfloat16, device=dev) for _ in range(4))
    out = torch.empty_like(q)
    lse = torch.empty(T, H, dtype=torch.float32, device=dev)
    qr = torch.tensor([[0, T]], dtype=torch.int32, device=dev)
    ty = torch.tensor([1], dtype=torch.int32, device=dev)
    for _ in range(2):
        _lib.ffa_fwd(q, k, v, out, lse, qr, qr, ty, D ** -0.5)
    torch.cuda.synchronize()
    lib = _lib.lib()
    lib.b200_debug_trace_fwd.argtypes = [C.c_void_p]
    lib.b200_debug_trace_bwd.argtypes = [C.c_void_p]
    dump("fwd (CTA 0 = heaviest q tile, kv blocks 64..127)", FWD, lib.b200_debug_trace_fwd)
    dq = torch.zeros(T, H, D, dtype=torch.float32, device=dev)
    dk, dv = torch.empty_like(k), torch.empty_like(v)
    for _ in range(2):
        _lib.ffa_bwd(do, q, k, v, out, lse, dq, dk, dv, qr, qr, ty, D ** -0.5)
    torch.cuda.synchronize()
    dump("bwd (CTA 0 = kv tile 0, q blocks 32..95)", BWD, lib.b200_debug_trace_bwd)


if __name__ == "__main__":
    main()
