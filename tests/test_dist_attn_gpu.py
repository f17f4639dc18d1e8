"""GPU tests of the context-parallel runtime.

* virtual-CP (1 GPU): every rank of a cp=4 plan is executed in turn on one device with the REAL kernels
  (stage accumulate forward, global-lse backward); the exchange itself is replaced by the oracle's naive
  world-level group_cast / group_reduce.  Checks planner slices + kernel accumulator semantics end to end.
* real CP (>=2 GPUs, else skipped): mp.spawn + NCCL through the public API (key / dispatch / calc_attn /
  undispatch / backward) against the global oracle.
"""
import os

import pytest
import torch

from mask_configs import mask_configs
from mp_util import spawn_with_deadline
from oracle import ffa_oracle as orc

pytestmark = pytest.mark.gpu
D = 128


def _slices(cfg):
    return [(q[0], q[1], k[0], k[1], t) for q, k, t in zip(cfg["q_ranges"], cfg["k_ranges"], cfg["types"])]


@pytest.mark.parametrize("name", ["causal", "sliding_window_causal", "varlen_block_causal_2k", "all_types_aspects"])
def test_virtual_cp_real_kernels(name):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from magiattention_b200.functional.dist_attn import DistAttnRuntime
    from magiattention_b200.meta.planner import make_attn_meta, make_dispatch_meta
    from magiattention_b200 import _lib

    dev = torch.device("cuda:0")
    cfg = next(c for c in mask_configs(1) if c["name"] == name)
    cp, chunk, degree, H, HK = 4, 128, 2, 2, 1
    T0 = cfg["seqlen"]
    T = T0 + (-T0) % (chunk * cp)
    sl = _slices(cfg)
    torch.manual_seed(7)
    q, k, v, do = (torch.randn(T, h, D, dtype=torch.bfloat16, device=dev) for h in (H, HK, HK, H))
    mask = orc.mask_from_slices([[s[0], s[1]] for s in sl], [[s[2], s[3]] for s in sl], [s[4] for s in sl], T, T)
    ro, rl = orc.attn_fwd(q, k, v, mask, high_precision=True)
    lo, ll = orc.attn_fwd(q, k, v, mask, high_precision=False)
    ref = orc.attn_bwd(do, q, k, v, ro.double(), rl.double(), mask, high_precision=True)
    lp = orc.attn_bwd(do, q, k, v, lo, ll, mask, high_precision=False)
    dm = make_dispatch_meta(sl, T, chunk, cp)
    comm, calc = make_attn_meta(sl, dm, degree)
    pos = [dm.position_ids(r).to(dev) for r in range(cp)]
    kl, vl = [k[p] for p in pos], [v[p] for p in pos]
    stage_k, stage_v = [], []
    for st in range(degree):
        args = [comm[r].kv_group_collective_args_list[st] for r in range(cp)]
        kw = dict(input_split_sizes=[a.input_split_size_list for a in args], dst_indices=[a.dst_indices_list for a in args],
                  src_index=[a.src_index_list for a in args], output_split_sizes=[a.output_split_size_list for a in args])
        stage_k.append(orc.group_cast_ref(kl, **kw))
        stage_v.append(orc.group_cast_ref(vl, **kw))
    scale = D ** -0.5
    out_g = torch.empty_like(q)
    lse_g = torch.empty(T, H, dtype=torch.float32, device=dev)
    dq_g = torch.empty(T, H, D, dtype=torch.float32, device=dev)
    dk_loc = [torch.zeros(dm.shard_seqlen, HK, D, dtype=torch.float32, device=dev) for _ in range(cp)]
    dv_loc = [torch.zeros(dm.shard_seqlen, HK, D, dtype=torch.float32, device=dev) for _ in range(cp)]
    parts = [[None] * cp for _ in range(degree)]
    for r in range(cp):
        rt = DistAttnRuntime(comm[r], calc[r], None)
        ql, dol = q[pos[r]].contiguous(), do[pos[r]].contiguous()
        out_acc = torch.zeros(ql.shape, dtype=torch.float32, device=dev)
        lse_acc = torch.full((ql.shape[0], H), float("-inf"), dtype=torch.float32, device=dev)
        rt.attn_fwd(ql, kl[r].contiguous(), vl[r].contiguous(), calc[r].local_attn_arg, out_acc, lse_acc, scale, 0.0)
        for st in range(degree):
            rt.attn_fwd(ql, stage_k[st][r].contiguous(), stage_v[st][r].contiguous(), calc[r].remote_attn_args_list[st],
                        out_acc, lse_acc, scale, 0.0)
        out_l = out_acc.to(torch.bfloat16)
        out_g[pos[r]] = out_l
        lse_g[pos[r]] = lse_acc
        dq = torch.zeros(ql.shape, dtype=torch.float32, device=dev)
        if not calc[r].local_attn_arg.can_skip:
            rt.attn_bwd(dol, ql, kl[r].contiguous(), vl[r].contiguous(), out_l, lse_acc, calc[r].local_attn_arg, dq, dk_loc[r],
                        dv_loc[r], scale, 0.0)
        for st in range(degree):
            rk, rv = stage_k[st][r].contiguous(), stage_v[st][r].contiguous()
            arg = calc[r].remote_attn_args_list[st]
            dkp = torch.zeros(rk.shape, dtype=torch.bfloat16, device=dev)
            dvp = torch.zeros(rv.shape, dtype=torch.bfloat16, device=dev)
            if not arg.can_skip:
                a = arg.to_ffa_args(dev)
                _lib.ffa_bwd(dol, ql, rk, rv, out_l, lse_acc, dq, dkp, dvp, a["q_ranges"], a["k_ranges"], a["attn_type_map"], scale)
            parts[st][r] = (dkp, dvp)
        dq_g[pos[r]] = dq
    for st in range(degree):
        rkw = [comm[r].kv_group_collective_args_list[st].to_group_reduce_args() for r in range(cp)]
        kw = dict(input_split_sizes=[a["input_split_sizes"] for a in rkw], dst_index=[a["dst_index"] for a in rkw],
                  src_indices=[a["src_indices"] for a in rkw], output_split_sizes=[a["output_split_sizes"] for a in rkw])
        dk_loc = orc.group_reduce_ref([parts[st][r][0] for r in range(cp)], dk_loc, **kw)
        dv_loc = orc.group_reduce_ref([parts[st][r][1] for r in range(cp)], dv_loc, **kw)
    dk_g = torch.empty(T, HK, D, dtype=torch.float32, device=dev)
    dv_g = torch.empty(T, HK, D, dtype=torch.float32, device=dev)
    for r in range(cp):
        dk_g[pos[r]] = dk_loc[r]
        dv_g[pos[r]] = dv_loc[r]
    orc.check_close("out", out_g, ro, lo, "out")
    orc.check_close("lse", lse_g, rl, ll, "lse")
    for nm, g, r_, l_ in zip(("dq", "dk", "dv"), (dq_g, dk_g, dv_g), ref, lp):
        orc.check_close(nm, g, r_, l_, nm, atol=orc.grad_atol(r_))


# ---------------------------------------------------------------------------------------------------
def _cp_worker(rank, world, port, name, with_sink=False, engine="nvl", iters=1, degree=None, hp_reduce=False):
    import torch.distributed as dist
    from magiattention_b200.api import calc_attn, compute_pad_size, dispatch, magi_attn_flex_key, undispatch
    from magiattention_b200.common import AttnRanges

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["MAGI_B200_KV_EXCHANGE"] = engine
    os.environ["MAGI_B200_NVL_TIMEOUT_S"] = "20"
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    import datetime
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=180))
    try:
        cfg = next(c for c in mask_configs(1) if c["name"] == name)
        T, H, HK, chunk = cfg["seqlen"], 4, 2, 128
        pad = compute_pad_size(T, world, chunk)
        torch.manual_seed(11)
        q, k, v, do = (torch.randn(T, h, D, dtype=torch.bfloat16, device=dev) for h in (H, HK, HK, H))
        for x in (q, k, v, do):
            dist.broadcast(x, 0)
        q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
        sink = None
        if with_sink:
            os.environ["MAGI_ATTENTION_DSINK_ALL_REDUCE_OP"] = "sum"
            sink = torch.randn(4, H, dtype=torch.float32, device=dev)
            dist.broadcast(sink, 0)
            sink.requires_grad_(True)
        from magiattention_b200.api import get_runtime_mgr
        from magiattention_b200.config import DistAttnConfig, OverlapConfig
        cfg_kw = {}
        if degree is not None or hp_reduce:
            cfg_kw["dist_attn_config"] = DistAttnConfig(overlap_config=OverlapConfig(degree=degree or 2), high_precision_reduce=hp_reduce)
        key = magi_attn_flex_key(AttnRanges.from_ranges(cfg["q_ranges"]), AttnRanges.from_ranges(cfg["k_ranges"]), cfg["types"],
                                 T, T, H, HK, D, pad, dist.group.WORLD, chunk_size=chunk, **cfg_kw)
        if world > 1:
            assert get_runtime_mgr(key).dist_attn_runtime.engine == engine, "requested exchange engine is not the one running"
        for it in range(iters):  # several passes: flag epochs advance, both arena slots get reused
            for x in (q, k, v) + ((sink,) if with_sink else ()):
                x.grad = None
            ql, kl, vl = dispatch(q, key), dispatch(k, key), dispatch(v, key)
            ol, meta = calc_attn(ql, kl, vl, key, sink=sink, return_max_logits=with_sink)
            out = undispatch(ol, key)
            out.backward(do)
        if world > 1 and engine == "nvl":
            from magiattention_b200.comm.nvl_exchange import NvlDomain
            torch.cuda.synchronize()
            NvlDomain.get(dist.group.WORLD).check()
        mask = orc.mask_from_slices(cfg["q_ranges"], cfg["k_ranges"], cfg["types"], T, T)
        qd, kd, vd = q.detach(), k.detach(), v.detach()
        sd = sink.detach() if with_sink else None
        ro, rl = orc.attn_fwd(qd, kd, vd, mask, high_precision=True, sink=sd)
        lo, ll = orc.attn_fwd(qd, kd, vd, mask, high_precision=False, sink=sd)
        orc.check_close("out", out.detach(), ro, lo, "out")
        if with_sink:
            torch.testing.assert_close(meta.max_logits.double(), orc.max_logits(qd.double(), kd.double(), mask), atol=1e-4, rtol=1e-5)
            ds_ref, ds_lp = orc.sink_bwd(sd, rl, ro, do, "sh"), orc.sink_bwd(sd, ll, lo, do, "sh")
            orc.check_close("dsink", sink.grad, ds_ref, ds_lp.float(), "dsink", atol=orc.dsink_atol(ds_ref),
                            min_mismatch=max(1.0 / ds_ref.numel(), 8e-2), norm_floor=0.015)
        ref = orc.attn_bwd(do, qd, kd, vd, ro.double(), rl.double(), mask, high_precision=True)
        lp = orc.attn_bwd(do, qd, kd, vd, lo, ll, mask, high_precision=False)
        for nm, g, r_, l_ in zip(("dq", "dk", "dv"), (q.grad, k.grad, v.grad), ref, lp):
            orc.check_close(nm, g, r_, l_, nm, atol=orc.grad_atol(r_))
    finally:
        from magiattention_b200.comm.nvl_exchange import NvlDomain
        NvlDomain.shutdown()
        dist.destroy_process_group()


@pytest.mark.parametrize("engine", ["nvl", "nccl"])
@pytest.mark.parametrize("name", ["causal", "sliding_window_causal"])
def test_real_cp_nccl(name, engine):
    """Both exchange engines (NVLink peer-memory pulls / NCCL P2P) through the public API against the global oracle."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(torch.cuda.device_count(), 4)
    spawn_with_deadline(_cp_worker, world, (name, False, engine), deadline_s=300)


@pytest.mark.parametrize("name,degree,hp", [("varlen_block_causal_2k", 3, False), ("all_types_aspects", 2, True),
                                            ("sparse_attn_2k", 4, False)])
def test_real_cp_nvl_multi_pass(name, degree, hp):
    """NVLink engine over several consecutive forward/backward passes (epoch flags advance, both arena slots are reused),
    more overlap stages than ranks need, and the fp32 wire dtype of the partial dK/dV."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(torch.cuda.device_count(), 4)
    spawn_with_deadline(_cp_worker, world, (name, False, "nvl", 4, degree, hp), deadline_s=300)


@pytest.mark.parametrize("name", ["uncovered_rows", "all_types_aspects"])
def test_real_cp_nccl_sink_and_max_logits(name):
    """calc_attn(sink=..., return_max_logits=True) through the runtime with the real kernels: sink applied once on the host
    stage, per-rank partial dsink all-reduced (MAGI_ATTENTION_DSINK_ALL_REDUCE_OP=sum), max_logits MAX-reduced.  Runs at
    cp = 1 on a single GPU (stage loop degenerate, same code path) and at cp = min(#GPUs, 4) otherwise."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    world = min(torch.cuda.device_count(), 4)
    spawn_with_deadline(_cp_worker, world, (name, True), deadline_s=300)
