"""world_size-2 (and 4) gloo tests, on CPU, of the N>1 exchange path: group_cast / group_reduce against the
naive world-level semantics, driven by the planner's real transfer tables, plus a full staged CP attention
(forward merge + dKV reduction) where the per-stage attention math is done by the oracle."""
import os

import pytest
import torch
import torch.distributed as dist

from mp_util import spawn_with_deadline
from oracle import ffa_oracle as orc


def _worker(rank, world, port, fn, args):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["MAGI_B200_ALLOW_HOST_COMM_TEST"] = "1"
    import datetime
    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=120))
    try:
        fn(rank, world, *args)
        dist.barrier()  # nobody tears the group down while a peer still has traffic in flight
    except BaseException:
        dist.destroy_process_group()
        raise
    dist.destroy_process_group()
    # skip interpreter teardown: gloo worker threads occasionally abort ("terminate called without an active
    # exception") when the autograd thread's last work handles are collected after the group is gone
    import sys
    sys.stdout.flush(); sys.stderr.flush()
    os._exit(0)


def _run(world, fn, *args):
    spawn_with_deadline(_worker, world, (fn, args))


def _plan(T, cp, chunk, degree, slices):
    from magiattention_b200.meta.planner import make_attn_meta, make_dispatch_meta

    dm = make_dispatch_meta(slices, T, chunk, cp)
    comm, calc = make_attn_meta(slices, dm, degree)
    return dm, comm, calc


def _cast_reduce_body(rank, world, T, chunk, degree, slices):
    from magiattention_b200.comm import group_cast, group_reduce

    dm, comm, calc = _plan(T, world, chunk, degree, slices)
    torch.manual_seed(0)
    kg = torch.randn(T, 2, 4)
    vg = torch.randn(T, 2, 4)
    pos = [dm.position_ids(r) for r in range(world)]
    kl, vl = [kg[p] for p in pos], [vg[p] for p in pos]
    for st in range(degree):
        args = [comm[r].kv_group_collective_args_list[st] for r in range(world)]
        kw = dict(input_split_sizes=[a.input_split_size_list for a in args], dst_indices=[a.dst_indices_list for a in args],
                  src_index=[a.src_index_list for a in args], output_split_sizes=[a.output_split_size_list for a in args])
        want_k = orc.group_cast_ref(kl, **kw)[rank]
        want_v = orc.group_cast_ref(vl, **kw)[rank]
        work = group_cast((kl[rank], vl[rank]), None, group=dist.group.WORLD, async_op=True, **args[rank].to_group_cast_args())
        rk, rv = work.wait_post_process()
        assert torch.equal(rk, want_k) and torch.equal(rv, want_v)
        # received rows really are the global rows the plan says (zero-redundant fetch of visible keys only)
        # transpose: reduce partial grads back
        parts = [torch.randn(sum(a.output_split_size_list), 2, 4, generator=torch.Generator().manual_seed(100 + r)) for r, a in
                 enumerate(args)]
        base = [torch.randn(dm.shard_seqlen, 2, 4, generator=torch.Generator().manual_seed(200 + r)) for r in range(world)]
        rkw = [a.to_group_reduce_args() for a in args]
        want = orc.group_reduce_ref(parts, base, input_split_sizes=[a["input_split_sizes"] for a in rkw],
                                    dst_index=[a["dst_index"] for a in rkw], src_indices=[a["src_indices"] for a in rkw],
                                    output_split_sizes=[a["output_split_sizes"] for a in rkw])[rank]
        out = base[rank].clone()
        w = group_reduce(parts[rank], out, group=dist.group.WORLD, async_op=True, **rkw[rank])
        w.wait_post_process()
        torch.testing.assert_close(out, want, atol=1e-6, rtol=1e-6)
        with pytest.raises(RuntimeError):
            w.wait_post_process()  # one-shot semantics (reference comm/work.py:83-114)


@pytest.mark.parametrize("world", [2, 4])
def test_group_cast_reduce_gloo(world):
    T = 512
    _run(world, _cast_reduce_body, T, 16, 2, [(0, T, 0, T, 1), ])


def test_group_cast_reduce_gloo_varlen():
    T = 384
    slices = [(0, 100, 0, 100, 1), (100, 228, 100, 228, 0), (228, 384, 130, 384, 3)]
    _run(2, _cast_reduce_body, T, 16, 3, slices)


def _dispatch_body(rank, world, T, chunk):
    from magiattention_b200.api import (calc_attn, compute_pad_size, dispatch, get_position_ids, magi_attn_flex_key,
                                        undispatch)
    from magiattention_b200.common import AttnRanges

    total = T - 5  # force padding
    pad = compute_pad_size(total, world, chunk)
    key = magi_attn_flex_key(AttnRanges.from_ranges([[0, total]]), AttnRanges.from_ranges([[0, total]]), "causal", total, total,
                             2, 1, 128, pad, dist.group.WORLD, chunk_size=chunk)
    torch.manual_seed(1)
    x = torch.randn(total, 3, requires_grad=True)
    xl = dispatch(x, key)
    assert xl.shape[0] == (total + pad) // world
    pos = get_position_ids(key).cpu()
    valid = pos < total
    assert torch.equal(xl[valid], x[pos[valid]])
    y = undispatch(xl * 2.0, key)
    assert torch.equal(y, x * 2.0)
    y.sum().backward()
    assert torch.equal(x.grad, torch.full_like(x, 2.0))
    # same key twice -> cache hit, same object
    key2 = magi_attn_flex_key(AttnRanges.from_ranges([[0, total]]), AttnRanges.from_ranges([[0, total]]), "causal", total, total,
                              2, 1, 128, pad, dist.group.WORLD, chunk_size=chunk)
    assert key2 == key
    # calc_attn on host tensors must fail loudly: the hot path has no CPU implementation
    q = torch.randn(xl.shape[0], 2, 128, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        calc_attn(q, q[:, :1], q[:, :1], key)


def test_dispatch_undispatch_gloo():
    _run(2, _dispatch_body, 256, 16)
