"""Reference arm of bench.py: the UNMODIFIED reference (MagiAttention, FFA_FA4 backend) timed on the same box.

`baseline/_ref` holds the reference built by tools/build_reference.sh (its own sources; nothing of this repo's kernels,
planner or runtime is imported here).  Single GPU: ``ffa_fa4_func(..., reuse_attn_arg=True)`` after one warm-up call that
builds the cached FA4AttnArg, exactly like the reference's own kernel benchmark
(exps/attn/run_blackwell_benchmark.py:859-885).  N > 1: ``magi_attn_flex_key`` / ``dispatch`` / ``calc_attn`` + backward
with the published CP tuning (exps/dist_attn/benchmark_conf.py:56-84,233-250: chunk 2048, static overlap degree 2, MinHeap
dispatch, CUDA_DEVICE_MAX_CONNECTIONS=8, NCCL_CGA_CLUSTER_SIZE=1, TORCH_NCCL_HIGH_PRIORITY=1, FA4 backend), key built outside
the timed region (exps/dist_attn/run_benchmark.py:591-624).

This module must run in a process that has NOT imported this repo's `magi_attention` alias package.
"""
from __future__ import annotations

import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> tuple[bool, str]:
    need = ["magi_attention", "flash_attn_cute"]
    missing = [n for n in need if not os.path.isdir(os.path.join(REF, n))]
    if missing:
        return False, f"baseline/_ref lacks {missing} (run tools/build_reference.sh)"
    if not any(f.startswith("magi_attn_ext") and f.endswith(".so") for f in os.listdir(os.path.join(REF, "magi_attention"))):
        return False, "baseline/_ref/magi_attention/magi_attn_ext*.so missing"
    for ext in ("magi_to_hstu_cuda", "create_block_mask_cuda"):
        if not any(f.startswith(ext) and f.endswith(".so") for f in os.listdir(REF)):
            return False, f"baseline/_ref/{ext}*.so missing"
    return True, ""


def setup_env(world: int) -> None:
    """The reference's published tuning, set before CUDA / NCCL initialise."""
    os.environ.setdefault("MAGI_ATTENTION_KERNEL_BACKEND", "fa4")
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "8")
    os.environ.setdefault("NCCL_CGA_CLUSTER_SIZE", "1")
    os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
    # CuTe-DSL JIT cache inside the tree so a second process on the same box reuses the compiled kernels
    os.environ.setdefault("CUTE_DSL_CACHE_DIR", os.path.join(REF, ".cute_cache"))
    sys.path[:0] = [REF, os.path.join(HERE, "stubs")]
    for m in list(sys.modules):
        if m == "magi_attention" or m.startswith("magi_attention."):
            raise RuntimeError("the reference arm needs a process that has not imported this repo's magi_attention alias")


def run(workload: dict, steps: int, warmup: int, rank: int, world: int, local_rank: int) -> dict:
    """Time the reference on ``workload`` (dict from bench.make_workload: slices, shapes).  Returns the fields of the
    JSON line that the caller completes.  Raises on any failure (the caller reports the reason and falls back)."""
    import torch
    import torch.distributed as dist

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    import magi_attention  # noqa: F401  (the reference)
    assert os.path.realpath(magi_attention.__file__).startswith(os.path.realpath(REF)), magi_attention.__file__
    from magi_attention.functional.fa4 import ffa_fa4_func, is_fa4_installed
    assert is_fa4_installed, "flash_attn_cute did not import"

    S, HQ, HK, D = workload["seq_len"], workload["num_heads_q"], workload["num_heads_kv"], workload["head_dim"]
    qr, kr, ty = workload["q_ranges"], workload["k_ranges"], workload["types"]
    g = torch.Generator(device="cpu").manual_seed(42)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / n
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    t_setup = time.time()
    if world == 1:
        q, k, v, do = (torch.randn(S, h, D, generator=g, dtype=torch.float32).to(torch.bfloat16).to(dev)
                       for h in (HQ, HK, HK, HQ))
        q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
        qrt = torch.tensor(qr, dtype=torch.int32, device=dev)
        krt = torch.tensor(kr, dtype=torch.int32, device=dev)
        tyt = torch.tensor(ty, dtype=torch.int32, device=dev)
        ffa_fa4_func(q, k, v, q_ranges=qrt, k_ranges=krt, attn_type_map=tyt, reuse_attn_arg=False)
        torch.cuda.synchronize()

        fwd_only = bool(workload.get("fwd_only"))

        def step():
            q.grad = k.grad = v.grad = None
            o, _ = ffa_fa4_func(q, k, v, q_ranges=qrt, k_ranges=krt, attn_type_map=tyt, reuse_attn_arg=True)
            if not fwd_only:
                o.backward(do)
        api = "ffa_fa4_func(reuse_attn_arg=True)"
    else:
        from magi_attention.api import calc_attn, compute_pad_size, dispatch, magi_attn_flex_key
        from magi_attention.common import AttnRanges
        from magi_attention.common.enum import AttnMaskType
        from magi_attention.config import DispatchConfig, DistAttnConfig, MinHeapDispatchAlg, OverlapConfig

        chunk = 2048
        pad = compute_pad_size(S, world, chunk)
        names = {0: "full", 1: "causal", 2: "inv_causal", 3: "bi_causal"}
        key = magi_attn_flex_key(
            q_ranges=AttnRanges.from_ranges(qr), k_ranges=AttnRanges.from_ranges(kr),
            attn_mask_type=[AttnMaskType(names[t]) for t in ty], total_seqlen_q=S, total_seqlen_k=S,
            num_heads_q=HQ, num_heads_kv=HK, head_dim=D, pad_size=pad, chunk_size=chunk, cp_group_or_mesh=dist.group.WORLD,
            dist_attn_config=DistAttnConfig(dispatch_config=DispatchConfig(alg=MinHeapDispatchAlg()),
                                            overlap_config=OverlapConfig(degree=2)))
        S_loc = (S + pad) // world
        # local shards drawn directly (i.i.d. data: statistically what dispatch() of a global tensor hands this rank)
        gl = torch.Generator(device="cpu").manual_seed(42 + 10 * rank)
        q, k, v, do = (torch.randn(S_loc, h, D, generator=gl, dtype=torch.float32).to(torch.bfloat16).to(dev)
                       for h in (HQ, HK, HK, HQ))
        q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)

        fwd_only = bool(workload.get("fwd_only"))

        def step():
            q.grad = k.grad = v.grad = None
            o, _ = calc_attn(q, k, v, key)
            if not fwd_only:
                o.backward(do)
        api = "magi_attn_flex_key / calc_attn (backend fa4, chunk 2048, degree 2, MinHeap)"
    for _ in range(max(warmup, 1)):
        step()
    torch.cuda.synchronize()
    setup_s = time.time() - t_setup
    ms = timed(step, steps)
    if os.environ.get("BENCH_REF_SPLIT") == "1" and world == 1 and not fwd_only:
        # developer aid: the reference's forward-only and backward-only times (its own kernel benchmark times them apart)
        def f_only():
            with torch.no_grad():
                ffa_fa4_func(q, k, v, q_ranges=qrt, k_ranges=krt, attn_type_map=tyt, reuse_attn_arg=True)
        ms_f = timed(f_only, steps)
        o, _ = ffa_fa4_func(q, k, v, q_ranges=qrt, k_ranges=krt, attn_type_map=tyt, reuse_attn_arg=True)

        def b_only():
            q.grad = k.grad = v.grad = None
            o.backward(do, retain_graph=True)
        b_only()
        ms_b = timed(b_only, steps)
        print(f"REF_SPLIT fwd {ms_f:.3f} ms  bwd {ms_b:.3f} ms  fwd+bwd {ms:.3f} ms", file=sys.stderr)
    return {"ms_per_step": ms, "api": api, "setup_s": setup_s}
