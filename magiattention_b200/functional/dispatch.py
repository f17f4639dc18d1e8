"""dispatch / undispatch of token-major tensors between the global order and this rank's permuted shard
(reference: magi_attention/functional/dispatch.py:24-136 — forward of dispatch is a local chunk+cat,
its backward an all-gather; undispatch is the mirror).  Outside the attention metric."""
from __future__ import annotations

import torch
import torch.distributed as dist

from ..meta.planner import DispatchMeta


def _take_chunks(x: torch.Tensor, meta: DispatchMeta, rank: int) -> torch.Tensor:
    cs = meta.chunk_size
    xv = x.reshape(meta.num_chunks, cs, *x.shape[1:])
    idx = torch.as_tensor(meta.partitions[rank], device=x.device)
    return xv.index_select(0, idx).reshape(len(meta.partitions[rank]) * cs, *x.shape[1:])


def _gather_global(x_local: torch.Tensor, meta: DispatchMeta, group) -> torch.Tensor:
    cs = meta.chunk_size
    if meta.cp_size == 1:
        gathered = x_local.unsqueeze(0)
    else:
        flat = x_local.new_empty((meta.cp_size * x_local.shape[0], *x_local.shape[1:]))
        dist.all_gather_into_tensor(flat, x_local.contiguous(), group=group)
        gathered = flat.reshape(meta.cp_size, *x_local.shape)
    out = x_local.new_empty((meta.num_chunks, cs, *x_local.shape[1:]))
    for r, part in enumerate(meta.partitions):
        idx = torch.as_tensor(part, device=x_local.device)
        out.index_copy_(0, idx, gathered[r].reshape(len(part), cs, *x_local.shape[1:]))
    return out.reshape(meta.total_seqlen, *x_local.shape[1:])


class _DispatchFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, meta, rank, group):
        ctx.meta, ctx.rank, ctx.group = meta, rank, group
        return _take_chunks(x, meta, rank)

    @staticmethod
    def backward(ctx, g):
        return _gather_global(g.contiguous(), ctx.meta, ctx.group), None, None, None


class _UndispatchFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_local, meta, rank, group, is_partial_grad):
        ctx.meta, ctx.rank, ctx.group, ctx.partial = meta, rank, group, is_partial_grad
        return _gather_global(x_local.contiguous(), meta, group)

    @staticmethod
    def backward(ctx, g):
        if ctx.partial and ctx.meta.cp_size > 1:
            # every rank holds a PARTIAL gradient of the global tensor: sum over ranks before taking the local chunks
            # (reference: undispatch_func(is_partial_grad=True) reduce-scatters, functional/dispatch.py:96-136)
            g = g.contiguous().clone()
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return _take_chunks(g, ctx.meta, ctx.rank), None, None, None, None


def dispatch_func(x_global: torch.Tensor, meta: DispatchMeta, rank: int, group) -> torch.Tensor:
    """Global [total_seqlen (padded), ...] -> this rank's permuted shard."""
    return _DispatchFunc.apply(x_global, meta, rank, group)


def undispatch_func(x_local: torch.Tensor, meta: DispatchMeta, rank: int, group, is_partial_grad: bool = False) -> torch.Tensor:
    """This rank's shard -> global [total_seqlen (padded), ...] (all-gather + unpermute)."""
    return _UndispatchFunc.apply(x_local, meta, rank, group, is_partial_grad)
