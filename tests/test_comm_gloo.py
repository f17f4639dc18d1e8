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


def _reduce_ops_body(rank, world, T, chunk, degree, slices):
    """reduce_op 'avg' / 'lse' and cast_lse against a naive world-level restatement
    (reference semantics: common/range_op/_range_reduce.py:152-357, grpcoll/_group_collective.py:81-408)."""
    from magiattention_b200.comm import group_cast, group_reduce

    dm, comm, calc = _plan(T, world, chunk, degree, slices)
    H, Dh = 2, 4
    for st in range(degree):
        args = [comm[r].kv_group_collective_args_list[st] for r in range(world)]
        gen = lambda seed, *shape: torch.randn(*shape, generator=torch.Generator().manual_seed(seed))
        # ---- cast_lse: (x, lse) rows travel together
        xs = [gen(10 + r, dm.shard_seqlen, H, Dh) for r in range(world)]
        ls = [gen(20 + r, dm.shard_seqlen, H) for r in range(world)]
        kw = dict(input_split_sizes=[a.input_split_size_list for a in args], dst_indices=[a.dst_indices_list for a in args],
                  src_index=[a.src_index_list for a in args], output_split_sizes=[a.output_split_size_list for a in args])
        want_x, want_l = orc.group_cast_ref(xs, **kw)[rank], orc.group_cast_ref(ls, **kw)[rank]
        gx, gl = group_cast(xs[rank], None, group=dist.group.WORLD, cast_lse=True, input_lse=ls[rank],
                            **args[rank].to_group_cast_args()).wait_post_process()
        assert torch.equal(gx, want_x) and torch.equal(gl, want_l)
        # ---- the transposed reduce with every op
        rkw = [a.to_group_reduce_args() for a in args]
        parts = [gen(100 + r, sum(a.output_split_size_list), H, Dh) for r, a in enumerate(args)]
        plse = [gen(300 + r, sum(a.output_split_size_list), H) for r, a in enumerate(args)]
        base = [gen(200 + r, dm.shard_seqlen, H, Dh) for r in range(world)]
        blse = [gen(400 + r, dm.shard_seqlen, H) for r in range(world)]
        # naive: walk every sender's input splits, deliver to the owner's output split in table order
        my = rkw[rank]
        contrib = [[] for _ in my["output_split_sizes"]]  # per output split: list of (rows, lse rows)
        for s in range(world):
            off, k_of = 0, {}
            for sz, d in zip(rkw[s]["input_split_sizes"], rkw[s]["dst_index"]):
                if sz > 0 and d == rank:
                    # the j-th split s sends to me lands in my j-th output split that lists s
                    j = k_of.get("n", 0)
                    outs_with_s = [i for i, (osz, srcs) in enumerate(zip(my["output_split_sizes"], my["src_indices"])) if osz > 0 and s in srcs]
                    contrib[outs_with_s[j]].append((parts[s][off:off + sz], plse[s][off:off + sz]))
                    k_of["n"] = j + 1
                off += sz
        for acc in (True, False):
            want_avg, want_o, want_l = base[rank].clone(), base[rank].clone(), blse[rank].clone()
            if not acc:
                want_avg.zero_(); want_o.zero_(); want_l.fill_(float("-inf"))
            o = 0
            for (sz, cs) in zip(my["output_split_sizes"], contrib):
                if sz > 0:
                    tot = want_avg[o:o + sz] + sum(c[0] for c in cs) if cs else want_avg[o:o + sz]
                    div = (len(cs) + 1) if acc else max(len(cs), 1)
                    want_avg[o:o + sz] = tot / div if (len(cs) + (1 if acc else 0)) > 1 else tot
                    for (x2, l2) in cs:
                        l1 = want_l[o:o + sz]
                        l = torch.logaddexp(l1, l2)
                        w1, w2 = torch.exp(l1 - l).nan_to_num(0.0), torch.exp(l2 - l).nan_to_num(0.0)
                        want_o[o:o + sz] = w1.unsqueeze(-1) * want_o[o:o + sz] + w2.unsqueeze(-1) * x2
                        want_l[o:o + sz] = l
                o += sz
            out = base[rank].clone()
            group_reduce(parts[rank], out, group=dist.group.WORLD, reduce_op="avg", acc_reduce=acc, **my).wait_post_process()
            torch.testing.assert_close(out, want_avg, atol=1e-5, rtol=1e-5)
            out, ol = base[rank].clone(), blse[rank].clone()
            group_reduce(parts[rank], out, group=dist.group.WORLD, reduce_op="lse", acc_reduce=acc, input_lse=plse[rank],
                         output_lse=ol, **my).wait_post_process()
            torch.testing.assert_close(out, want_o, atol=1e-5, rtol=1e-5)
            torch.testing.assert_close(ol, want_l, atol=1e-5, rtol=1e-5)
    with pytest.raises(ValueError):
        group_reduce(parts[rank], out, group=dist.group.WORLD, reduce_op="max", **my)


@pytest.mark.parametrize("world", [2, 4])
def test_group_reduce_avg_lse_and_cast_lse_gloo(world):
    T = 512
    _run(world, _reduce_ops_body, T, 16, 2, [(0, T, 0, T, 1), ])


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
