"""Step-by-step bring-up probe of the NVLink peer-memory exchange primitives (run under torchrun, >= 2 ranks):
arena alloc / IPC export+import, local flag signal+wait, remote flag signal, copy-engine pulls from a peer, the wait
timeout, then one forward and one backward pass of the runtime with per-step flag dumps."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def log(rank, *a):
    print(f"[r{rank} {time.time() % 1000:7.2f}]", *a, flush=True)


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    from magiattention_b200 import _lib
    from magiattention_b200.comm.nvl_exchange import _RawCuda

    peer = (rank + 1) % world
    log(rank, "can_access_peer", [torch.cuda.can_device_access_peer(lr, d) for d in range(world) if d != lr])
    nbytes = 64 << 20
    base = _lib.nvl_alloc(nbytes)
    handles = [None] * world
    dist.all_gather_object(handles, _lib.nvl_export(base))
    pb = [base if r == rank else _lib.nvl_import(h) for r, h in enumerate(handles)]
    log(rank, "arena", hex(base), "peers", [hex(x) for x in pb])
    view = torch.as_tensor(_RawCuda(base, nbytes, None), device=dev)
    assert view.data_ptr() == base
    flags = view[:16384].view(torch.int32)
    err = torch.zeros(4, dtype=torch.int32).pin_memory()
    s1 = torch.cuda.Stream(device=dev, priority=-1)
    T3 = int(3e9)

    # 1. local signal + wait
    f_loc = _lib.NvlFlagList([base + 4 * 1000])
    f_loc.signal(5, s1.cuda_stream)
    f_loc.wait(5, err.data_ptr(), T3, s1.cuda_stream)
    s1.synchronize()
    log(rank, "1 local signal/wait: flag", int(flags[1000]), "err", int(err[0]))
    # 2. wait timeout on a flag nobody sets
    t0 = time.time()
    _lib.NvlFlagList([base + 4 * 1001]).wait(1, err.data_ptr(), int(1e9), s1.cuda_stream)
    s1.synchronize()
    log(rank, f"2 timeout wait returned after {time.time() - t0:.2f}s err", int(err[0]))
    err.zero_()
    dist.barrier()
    # 3. remote signal: write 7 into word 2000+rank of every peer, wait for every peer's word here
    _lib.NvlFlagList([pb[p] + 4 * (2000 + rank) for p in range(world) if p != rank]).signal(7, s1.cuda_stream)
    _lib.NvlFlagList([base + 4 * (2000 + p) for p in range(world) if p != rank]).wait(7, err.data_ptr(), T3, s1.cuda_stream)
    s1.synchronize()
    log(rank, "3 remote signal: my flags", flags[2000:2000 + world].tolist(), "err", int(err[0]))
    err.zero_()
    dist.barrier()
    # 4. copy-engine pull from the peer's arena
    data = view[1 << 20:(1 << 20) + (32 << 20)].view(torch.int32)
    data.fill_(rank + 1)
    torch.cuda.synchronize()
    dist.barrier()
    dst = torch.zeros(8 << 20, dtype=torch.int32, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cl = _lib.NvlCopyList([(dst.data_ptr(), pb[peer] + (1 << 20), 32 << 20)])
    with torch.cuda.stream(s1):
        e0.record()
        cl.launch(s1.cuda_stream)
        e1.record()
    s1.synchronize()
    log(rank, "4 pull 32 MiB from peer", peer, "ok", bool((dst == peer + 1).all()), f"{(32 << 20) / e0.elapsed_time(e1) / 1e6:.1f} GB/s")
    big = 1 << 30
    if nbytes < big + (2 << 20):
        pass
    dist.barrier()
    for p in pb:
        if p != base:
            _lib.nvl_release(p)
    del view, flags, data
    torch.cuda.synchronize()
    _lib.nvl_free(base)
    dist.barrier()

    # 5. the runtime: forward, then backward, with the domain's flags dumped after each
    from magiattention_b200.api import calc_attn, get_runtime_mgr, magi_attn_flex_key
    from magiattention_b200.common import AttnRanges
    from magiattention_b200.comm.nvl_exchange import NvlDomain
    from magiattention_b200.config import DistAttnConfig, OverlapConfig

    os.environ["MAGI_B200_NVL_TIMEOUT_S"] = "5"
    NvlDomain.TIMEOUT_NS = int(5e9)
    S, H, D = 8192 * world, 4, 128
    r_ = AttnRanges.from_ranges([[0, S]])
    key = magi_attn_flex_key(r_, r_, "causal", S, S, H, H, D, 0, dist.group.WORLD,
                             dist_attn_config=DistAttnConfig(overlap_config=OverlapConfig(degree=2)), chunk_size=512)
    rt = get_runtime_mgr(key).dist_attn_runtime
    log(rank, "5 engine", rt.engine, "pull plan", [[(p[0], p[3]) for p in pl] for pl in rt.nvl.pull] if rt.nvl else None)
    q, k, v, do = (torch.randn(S // world, H, D, dtype=torch.bfloat16, device=dev) for _ in range(4))
    q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)

    def dump(tag):
        torch.cuda.synchronize()
        dom = rt.nvl.dom
        fl = dom.view[:16384].view(torch.int32)
        log(rank, tag, "pass", dom.pass_id, "kv_ready", fl[0:world].tolist(), "pulls_done", fl[64:64 + world].tolist(),
            "part_ready", [fl[128 + s * 32:128 + s * 32 + 2].tolist() for s in range(world)], "err", int(dom.err[0]))

    for it in range(3):
        q.grad = k.grad = v.grad = None
        o, _ = calc_attn(q, k, v, key)
        dump(f"6.{it} after fwd")
        o.backward(do)
        dump(f"7.{it} after bwd")
    dist.barrier()
    NvlDomain.shutdown()
    dist.destroy_process_group()
    log(rank, "done")


if __name__ == "__main__":
    main()
