"""Cyclic roll of a dispatched tensor along the (global, padded) sequence — for multi-token-prediction labels.

Same contract as the reference (magi_attention/functional/roll.py; api roll / roll_simple, magi_attn_interface.py:960-1038):
the result equals ``dispatch(torch.roll(undispatch(x), shift))`` over the dispatch's ``total_seqlen`` rows, without ever
materialising the global tensor.  Engine: every rank derives, from the DispatchMeta alone, the maximal row spans
``(source rank, source local rows) -> (my local rows)`` of the rolled layout; spans that stay on the rank are device copies,
the others are one grouped P2P batch (``roll``) or plain isend/irecv pairs (``roll_simple``) straight between the row spans.
Backward is the roll by ``-shift``.  Outside the attention metric (SURVEY §8f-1 neighbourhood).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from ..meta.planner import DispatchMeta

__all__ = ["roll_func", "roll_simple_func", "roll_spans"]


def roll_spans(meta: DispatchMeta, shift: int) -> list[list[tuple[int, int, int, int]]]:
    """For every destination rank: ``(src_rank, src_local_start, dst_local_start, length)`` spans, in destination order.
    A span never crosses a chunk boundary on either side, adjacent spans that are contiguous on both sides are merged."""
    cs, n, T = meta.chunk_size, meta.num_chunks, meta.total_seqlen
    owner, loc = [0] * n, [0] * n
    for r, part in enumerate(meta.partitions):
        for i, ch in enumerate(part):
            owner[ch], loc[ch] = r, i * cs
    shift %= T
    plans: list[list[tuple[int, int, int, int]]] = []
    for r, part in enumerate(meta.partitions):
        spans: list[tuple[int, int, int, int]] = []
        for i, ch in enumerate(part):
            g, dst, left = (ch * cs - shift) % T, i * cs, cs  # out[global p] = in[(p - shift) mod T]
            while left > 0:
                sc = g // cs
                take = min(left, (sc + 1) * cs - g)
                s_rank, s_loc = owner[sc], loc[sc] + (g - sc * cs)
                if spans and spans[-1][0] == s_rank and spans[-1][1] + spans[-1][3] == s_loc and spans[-1][2] + spans[-1][3] == dst:
                    spans[-1] = (s_rank, spans[-1][1], spans[-1][2], spans[-1][3] + take)
                else:
                    spans.append((s_rank, s_loc, dst, take))
                g, dst, left = (g + take) % T, dst + take, left - take
        plans.append(spans)
    return plans


def _global_rank(group, r: int) -> int:
    return r if group is None or group is dist.group.WORLD else dist.get_global_rank(group, r)


def _roll_impl(x_local: torch.Tensor, shift: int, meta: DispatchMeta, rank: int, group, seq_dim: int, batched: bool) -> torch.Tensor:
    if shift % meta.total_seqlen == 0:
        return x_local.clone()
    x = x_local.movedim(seq_dim, 0).contiguous()
    assert x.shape[0] == meta.shard_seqlen, f"expected {meta.shard_seqlen} local rows along dim {seq_dim}, got {x.shape[0]}"
    out = torch.empty_like(x)
    plans = roll_spans(meta, shift)
    ops, pending = [], []
    # both sides enumerate the same plan, so the k-th message between a pair of ranks is the same span on both ends
    recvs: dict[int, list] = {}
    sends: dict[int, list] = {}
    for (src, s0, d0, n) in plans[rank]:
        if src == rank:
            out[d0:d0 + n].copy_(x[s0:s0 + n])
        else:
            recvs.setdefault(src, []).append(out[d0:d0 + n])
    for dst_rank, spans in enumerate(plans):
        if dst_rank != rank:
            for (src, s0, _, n) in spans:
                if src == rank:
                    sends.setdefault(dst_rank, []).append(x[s0:s0 + n])
    for peer in sorted(set(recvs) | set(sends)):
        gp = _global_rank(group, peer)
        if batched:
            ops += [dist.P2POp(dist.irecv, t, gp, group) for t in recvs.get(peer, [])]
            ops += [dist.P2POp(dist.isend, t, gp, group) for t in sends.get(peer, [])]
            continue
        # un-batched NCCL P2P ops of one rank run in issue order on one stream: two ranks that both post their receive
        # first wait for each other forever.  Matched pair order (reference functional/roll.py:285-350): the lower rank
        # sends first, the higher rank receives first; peers are visited in increasing order on every rank.
        first, second = (sends, recvs) if rank < peer else (recvs, sends)
        for table in (first, second):
            fn = dist.isend if table is sends else dist.irecv
            pending += [fn(t, gp, group=group) for t in table.get(peer, [])]
    if ops:
        pending = dist.batch_isend_irecv(ops)
    for w in pending:
        w.wait()
    return out.movedim(0, seq_dim)


class _Roll(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_local, shift, meta, rank, group, seq_dim, batched):
        ctx.args = (shift, meta, rank, group, seq_dim, batched)
        return _roll_impl(x_local, shift, meta, rank, group, seq_dim, batched)

    @staticmethod
    def backward(ctx, g):
        shift, meta, rank, group, seq_dim, batched = ctx.args
        return _roll_impl(g, -shift, meta, rank, group, seq_dim, batched), None, None, None, None, None, None


def roll_func(x_local: torch.Tensor, shift: int, meta: DispatchMeta, rank: int, group, seq_dim: int = 0) -> torch.Tensor:
    """Grouped-P2P roll (one ``batch_isend_irecv``)."""
    return _Roll.apply(x_local, shift, meta, rank, group, seq_dim, True)


def roll_simple_func(x_local: torch.Tensor, shift: int, meta: DispatchMeta, rank: int, group, seq_dim: int = 0) -> torch.Tensor:
    """Same result with plain isend / irecv pairs."""
    return _Roll.apply(x_local, shift, meta, rank, group, seq_dim, False)
