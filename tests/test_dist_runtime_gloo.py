"""CPU (gloo, world 2 and 4) test of the WHOLE context-parallel runtime — key / planner / dispatch / DistAttnFunc stage
loop / group_cast / group_reduce / undispatch / autograd — with the CUDA kernels replaced, inside this test only,
by the oracle acting as an exact fake kernel with the same accumulator semantics (the reference tests its runtime the
same way with its SDPA backends, tests/README.md:5-8).  The product itself has no such path."""
import os

import pytest
import torch
import torch.distributed as dist

from mask_configs import mask_configs
from oracle import ffa_oracle as orc
from test_comm_gloo import _run


def _install_fake_kernels():
    import magiattention_b200._lib as _lib
    import magiattention_b200.functional.dist_attn as da

    def _mask(qr, kr, ty, tq, tk):
        return orc.mask_from_slices(qr.tolist(), kr.tolist(), ty.tolist(), tq, tk)

    def fake_fwd(q, k, v, sink, sink_layout, out, lse, qr, kr, ty, scale, softcap, out_type, **kw):
        mask = _mask(qr, kr, ty, q.shape[0], k.shape[0])
        o, l = orc.attn_fwd(q.double(), k.double(), v.double(), mask, scale, sink=sink, sink_layout=sink_layout)
        mo, ml = orc.merge_out_lse(out.double(), lse.double(), o, l.double())
        out.copy_(mo)
        lse.copy_(ml)
        if kw.get("return_max_logits"):
            kw["max_logits"].copy_(torch.maximum(kw["max_logits"], orc.max_logits(q.double(), k.double(), mask, scale).float()))
        return out, None

    def _bwd(do, q, k, v, out, lse, qr, kr, ty, scale):
        return orc.attn_bwd(do.double(), q.double(), k.double(), v.double(), out.double(), lse.double(),
                            _mask(qr, kr, ty, q.shape[0], k.shape[0]), scale)

    def fake_bwd(do, q, k, v, sink, sink_layout, out, lse, dq, dk, dv, dsink, qr, kr, ty, scale, softcap, **kw):
        a, b, c = _bwd(do, q, k, v, out, lse, qr, kr, ty, scale)
        dq += a.to(dq.dtype); dk += b.to(dk.dtype); dv += c.to(dv.dtype)
        if sink is not None:
            dsink += orc.sink_bwd(sink, lse, out, do, sink_layout).to(dsink.dtype)
        return dq, dk, dv, dsink

    def fake_lib_bwd(do, q, k, v, out, lse, dq, dk, dv, qr, kr, ty, scale, softcap=0.0, accumulate_dkv=False, sm_margin=0, row_stats=None):
        a, b, c = _bwd(do, q, k, v, out, lse, qr, kr, ty, scale)
        dq += a.to(dq.dtype)
        dk.copy_(b); dv.copy_(c)
        return 3

    da._flex_flash_attn_forward = fake_fwd
    da._flex_flash_attn_backward = fake_bwd
    _lib.ffa_bwd = fake_lib_bwd


def _runtime_body(rank, world, name, degree, alg=None, hp_reduce=False):
    from magiattention_b200.api import calc_attn, compute_pad_size, dispatch, magi_attn_flex_key, undispatch
    from magiattention_b200.common import AttnRanges
    from magiattention_b200.common.enum import AttnOverlapMode
    from magiattention_b200.config import DispatchConfig, DistAttnConfig, OverlapConfig

    _install_fake_kernels()
    cfg = next(c for c in mask_configs(4) if c["name"] == name)
    T, H, HK, D, chunk = cfg["seqlen"], 2, 1, 8, 16
    pad = compute_pad_size(T, world, chunk)
    torch.manual_seed(21)
    q, k, v, do = (torch.randn(T, h, D, dtype=torch.float64) for h in (H, HK, HK, H))
    for x in (q, k, v, do):
        dist.broadcast(x, 0)
    q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
    key = magi_attn_flex_key(AttnRanges.from_ranges(cfg["q_ranges"]), AttnRanges.from_ranges(cfg["k_ranges"]), cfg["types"], T, T,
                             H, HK, D, pad, dist.group.WORLD,
                             dist_attn_config=DistAttnConfig(
                                 dispatch_config=DispatchConfig(alg=alg) if alg is not None else DispatchConfig(),
                                 overlap_config=(OverlapConfig(degree=degree) if degree is not None else
                                                 OverlapConfig(mode=AttnOverlapMode.DYNAMIC, degree=None)),
                                 high_precision_reduce=hp_reduce), chunk_size=chunk)
    ol, meta = calc_attn(dispatch(q, key), dispatch(k, key), dispatch(v, key), key)
    out = undispatch(ol, key)
    out.backward(do)
    mask = orc.mask_from_slices(cfg["q_ranges"], cfg["k_ranges"], cfg["types"], T, T)
    ro, rl = orc.attn_fwd(q.detach(), k.detach(), v.detach(), mask)
    rdq, rdk, rdv = orc.attn_bwd(do, q.detach(), k.detach(), v.detach(), ro, rl, mask)
    torch.testing.assert_close(out.detach(), ro, atol=1e-5, rtol=1e-5)  # fp32 stage accumulators
    lse_g = undispatch(meta.lse, key)
    fin = ~torch.isinf(rl)
    assert torch.equal(torch.isinf(lse_g), ~fin)
    torch.testing.assert_close(lse_g[fin].double(), rl[fin].double(), atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(q.grad, rdq, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(k.grad, rdk, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(v.grad, rdv, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("world,degree", [(2, 1), (4, 2), (4, 3)])
@pytest.mark.parametrize("name", ["causal", "sliding_window_causal", "sparse_attn_2k"])
def test_cp_runtime_end_to_end_gloo(world, degree, name):
    _run(world, _runtime_body, name, degree)


def _sink_body(rank, world, name, degree):
    """Sink applied once (host stage), partial dsink per rank summed over the group, max_logits MAX-reduced."""
    from magiattention_b200.api import calc_attn, compute_pad_size, dispatch, magi_attn_flex_key, undispatch
    from magiattention_b200.common import AttnRanges
    from magiattention_b200.config import DistAttnConfig, OverlapConfig

    os.environ["MAGI_ATTENTION_DSINK_ALL_REDUCE_OP"] = "sum"
    _install_fake_kernels()
    cfg = next(c for c in mask_configs(4) if c["name"] == name)
    T, H, HK, D, chunk = cfg["seqlen"], 2, 1, 8, 16
    pad = compute_pad_size(T, world, chunk)
    torch.manual_seed(33)
    q, k, v, do = (torch.randn(T, h, D, dtype=torch.float64) for h in (H, HK, HK, H))
    sink = torch.randn(3, H, dtype=torch.float64)
    for x in (q, k, v, do, sink):
        dist.broadcast(x, 0)
    for x in (q, k, v, sink):
        x.requires_grad_(True)
    key = magi_attn_flex_key(AttnRanges.from_ranges(cfg["q_ranges"]), AttnRanges.from_ranges(cfg["k_ranges"]), cfg["types"], T, T,
                             H, HK, D, pad, dist.group.WORLD,
                             dist_attn_config=DistAttnConfig(overlap_config=OverlapConfig(degree=degree)), chunk_size=chunk)
    ol, meta = calc_attn(dispatch(q, key), dispatch(k, key), dispatch(v, key), key, sink=sink, return_max_logits=True)
    out = undispatch(ol, key)
    out.backward(do)
    mask = orc.mask_from_slices(cfg["q_ranges"], cfg["k_ranges"], cfg["types"], T, T)
    qd, kd, vd, sd = (x.detach() for x in (q, k, v, sink))
    ro, rl = orc.attn_fwd(qd, kd, vd, mask, sink=sd, sink_layout="sh")
    rdq, rdk, rdv = orc.attn_bwd(do, qd, kd, vd, ro, rl, mask)
    torch.testing.assert_close(out.detach(), ro, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(undispatch(meta.lse, key).double(), rl.double(), atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(meta.max_logits.double(), orc.max_logits(qd, kd, mask), atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(q.grad, rdq, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(k.grad, rdk, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(v.grad, rdv, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(sink.grad, orc.sink_bwd(sd, rl, ro, do, "sh"), atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("world,degree,name", [(2, 1, "causal"), (4, 2, "uncovered_rows")])
def test_cp_runtime_sink_and_max_logits_gloo(world, degree, name):
    _run(world, _sink_body, name, degree)


def _new_mask_body(rank, world, degree):
    """Hybrid attention: dispatch once with the causal key, then run a sliding-window mask and a block-sparse mask on the
    SAME local shards through make_flex_key_for_new_mask_after_dispatch (reference api/magi_attn_interface.py:1315-1530)."""
    from magiattention_b200.api import (calc_attn, compute_pad_size, dispatch, get_position_ids, magi_attn_flex_key,
                                        make_flex_key_for_new_mask_after_dispatch, undispatch)
    from magiattention_b200.common import AttnRanges
    from magiattention_b200.config import DistAttnConfig, OverlapConfig

    _install_fake_kernels()
    cfgs = {c["name"]: c for c in mask_configs(4)}
    base = cfgs["causal"]
    T, H, HK, D, chunk = base["seqlen"], 2, 1, 8, 16
    pad = compute_pad_size(T, world, chunk)
    torch.manual_seed(5)
    q, k, v, do = (torch.randn(T, h, D, dtype=torch.float64) for h in (H, HK, HK, H))
    for x in (q, k, v, do):
        dist.broadcast(x, 0)
    key0 = magi_attn_flex_key(AttnRanges.from_ranges(base["q_ranges"]), AttnRanges.from_ranges(base["k_ranges"]), base["types"],
                              T, T, H, HK, D, pad, dist.group.WORLD,
                              dist_attn_config=DistAttnConfig(overlap_config=OverlapConfig(degree=degree)), chunk_size=chunk)
    ql, kl, vl = (dispatch(x, key0).requires_grad_(True) for x in (q, k, v))
    for name in ("sliding_window_causal", "sparse_attn_2k"):
        cfg = cfgs[name]
        assert cfg["seqlen"] == T
        key1 = make_flex_key_for_new_mask_after_dispatch(AttnRanges.from_ranges(cfg["q_ranges"]),
                                                         AttnRanges.from_ranges(cfg["k_ranges"]), cfg["types"], key0)
        assert key1 != key0 and torch.equal(get_position_ids(key1), get_position_ids(key0))  # same dispatch solution
        ol, _ = calc_attn(ql, kl, vl, key1)
        out = undispatch(ol, key1)
        mask = orc.mask_from_slices(cfg["q_ranges"], cfg["k_ranges"], cfg["types"], T, T)
        ro, rl = orc.attn_fwd(q, k, v, mask)
        torch.testing.assert_close(out.detach(), ro, atol=1e-5, rtol=1e-5)
        for x in (ql, kl, vl):
            x.grad = None
        out.backward(do)
        rdq, rdk, rdv = orc.attn_bwd(do, q, k, v, ro, rl, mask)
        for g, r in ((ql.grad, rdq), (kl.grad, rdk), (vl.grad, rdv)):
            torch.testing.assert_close(undispatch(g, key0), r, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("world,degree", [(2, 1), (4, 2)])
def test_new_mask_after_dispatch_gloo(world, degree):
    _run(world, _new_mask_body, degree)


@pytest.mark.parametrize("alg_name", ["sequential", "sorted_sequential"])
def test_cp_runtime_other_dispatch_algorithms_gloo(alg_name):
    """The runtime result does not depend on the dispatch solution: contiguous (ring-style) and snake dispatch, degree None
    (planner-chosen stage count)."""
    from magiattention_b200.config import SequentialDispatchAlg, SortedSequentialSelectAlg

    alg = SequentialDispatchAlg() if alg_name == "sequential" else SortedSequentialSelectAlg()
    _run(4, _runtime_body, "all_types_aspects", None, alg)


def _roll_body(rank, world):
    """roll / roll_simple == dispatch(torch.roll(undispatch(x))) incl. padding rows, any dim, autograd (= roll by -shift)."""
    from magiattention_b200.api import compute_pad_size, dispatch, magi_attn_flex_key, roll, roll_simple, undispatch
    from magiattention_b200.api.magi_attn_interface import get_runtime_mgr
    from magiattention_b200.common import AttnRanges

    T, chunk = 200, 16
    pad = compute_pad_size(T, world, chunk)
    key = magi_attn_flex_key(AttnRanges.from_ranges([[0, T]]), AttnRanges.from_ranges([[0, T]]), "causal", T, T, 2, 2, 8, pad,
                             dist.group.WORLD, chunk_size=chunk)
    mgr = get_runtime_mgr(key)
    Tp, chunk = key.total_seqlen_q, key.chunk_size  # padded total / the chunk size the key settled on (capped like the reference)
    torch.manual_seed(3)
    g = torch.randn(Tp, 3, dtype=torch.float64)  # the padded global tensor as the dispatch sees it
    dist.broadcast(g, 0)
    pos = mgr.get_position_ids()
    x = g[pos].clone().requires_grad_(True)
    for fn in (roll, roll_simple):
        for shift in (1, -1, 7, chunk, chunk + 3, -45, Tp, Tp + 2):
            y = fn(x, shift, 0, key)
            assert torch.equal(y.detach(), torch.roll(g, shift, 0)[pos]), (fn.__name__, shift)
    # a non-leading sequence dim and the backward pass
    xt = x.detach().t().contiguous().requires_grad_(True)  # [3, rows]
    yt = roll(xt, 5, 1, key)
    assert torch.equal(yt.detach(), torch.roll(g, 5, 0)[pos].t())
    w = torch.randn(Tp, 3, dtype=torch.float64)
    dist.broadcast(w, 0)
    (yt * w[pos].t()).sum().backward()
    assert torch.equal(xt.grad, torch.roll(w, -5, 0)[pos].t())


@pytest.mark.parametrize("world", [2, 4])
def test_roll_gloo(world):
    _run(world, _roll_body)


def test_cp_runtime_fp32_wire_for_dkv_gloo():
    """high_precision_reduce: the partial dK/dV of every stage travel as fp32 (reference env switch
    MAGI_ATTENTION_BACKWARD_HIGH_PRECISION_REDUCE / our DistAttnConfig field)."""
    _run(4, _runtime_body, "sliding_window_causal", 2, None, True)
