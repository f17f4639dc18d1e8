"""One rank's share of BASELINE config C5 (1M tokens, 64:8 GQA, block-sparse rectangles, cp=8) on ONE GPU: the real planner,
the real kernels stage by stage (remote K/V = random stage buffers of the planned sizes: timing only, parity is
tests/test_named_configs_gpu.py), per-kernel times from the torch profiler — in particular plan_kernel vs the attention kernels."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT]
import torch  # noqa: E402

from bench_workloads import make_workload  # noqa: E402
from magiattention_b200 import _lib  # noqa: E402
from magiattention_b200.meta.planner import make_attn_meta, make_dispatch_meta  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    w = make_workload("C5")
    S, HQ, HK, D, cp, degree, rank = w["seq_len"], 64, 8, 128, 8, 4, 3
    sl = [(q[0], q[1], k[0], k[1], t) for q, k, t in zip(w["q_ranges"], w["k_ranges"], w["types"])]
    dm = make_dispatch_meta(sl, S, 2048, cp)
    comm, calc = make_attn_meta(sl, dm, degree)
    c = calc[rank]
    args = [c.local_attn_arg] + list(c.remote_attn_args_list)
    rows = [dm.shard_seqlen] + list(c.seqlen_k_remote_per_stage)
    area = sum(a.total_area for a in args)
    torch.manual_seed(1)
    q, do = (torch.randn(dm.shard_seqlen, HQ, D, dtype=torch.bfloat16, device=dev) for _ in range(2))
    kv = [(torch.randn(n, HK, D, dtype=torch.bfloat16, device=dev), torch.randn(n, HK, D, dtype=torch.bfloat16, device=dev)) for n in rows]
    out = torch.zeros(q.shape, dtype=torch.float32, device=dev)
    lse = torch.full((q.shape[0], HQ), float("-inf"), dtype=torch.float32, device=dev)
    scale = D ** -0.5

    def step():
        out.zero_(); lse.fill_(float("-inf"))
        for a, (k, v) in zip(args, kv):
            f = a.to_ffa_args(dev)
            _lib.ffa_fwd(q, k, v, out, lse, f["q_ranges"], f["k_ranges"], f["attn_type_map"], scale, accumulate=True)
        ob = out.to(torch.bfloat16)
        dq = torch.zeros(q.shape, dtype=torch.float32, device=dev)
        for a, (k, v) in zip(args, kv):
            f = a.to_ffa_args(dev)
            dk, dv = torch.empty_like(k), torch.empty_like(v)
            _lib.ffa_bwd(do, q, k, v, ob, lse, dq, dk, dv, f["q_ranges"], f["k_ranges"], f["attn_type_map"], scale)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    per = {}
    for ev in prof.key_averages():
        t = getattr(ev, "device_time_total", None) or getattr(ev, "cuda_time_total", 0)
        name = ev.key.split("<")[0].split("(")[0]
        per[name] = per.get(name, 0.0) + t / 1e3
    tot = sum(per.values())
    res = {"workload": "C5 one rank's share (rank 3 of cp=8): 128K local q rows x 64/8 heads, local + 4 remote stages",
           "slices_per_call": [len(a.q_ranges) for a in args], "kv_rows_per_call": rows, "area": area,
           "ms_per_step_fwd_bwd": ms, "tflops": 3.5 * 4.0 * area * HQ * D / ms / 1e9,
           "kernel_ms": {k: round(v, 3) for k, v in sorted(per.items(), key=lambda x: -x[1])[:8]},
           "plan_kernel_pct_of_step": 100.0 * sum(v for k, v in per.items() if "plan" in k) / max(tot, 1e-9)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
