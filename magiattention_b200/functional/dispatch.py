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
    def forward(ctx, x, meta, rank, group, pad):
        ctx.meta, ctx.rank, ctx.group, ctx.pad = meta, rank, group, pad
        if pad:
            x = torch.cat([x, x.new_zeros((pad, *x.shape[1:]))], 0)
        return _take_chunks(x, meta, rank)

    @staticmethod
    def backward(ctx, g):
        full = _gather_global(g.contiguous(), ctx.meta, ctx.group)
        if ctx.pad:
            full = full[: full.shape[0] - ctx.pad]
        return full, None, None, None, None


class _UndispatchFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_local, meta, rank, group, pad):
        ctx.meta, ctx.rank, ctx.group, ctx.pad = meta, rank, group, pad
        full = _gather_global(x_local.contiguous(), meta, group)
        return full[: full.shape[0] - pad] if pad else full

    @staticmethod
    def backward(ctx, g):
        if ctx.pad:
            g = torch.cat([g, g.new_zeros((ctx.pad, *g.shape[1:]))], 0)
        return _take_chunks(g, ctx.meta, ctx.rank), None, None, None, None


def dispatch_func(x, meta: DispatchMeta, rank: int, group, pad_size: int = 0) -> torch.Tensor:
    return _DispatchFunc.apply(x, meta, rank, group, pad_size)


def undispatch_func(x_local, meta: DispatchMeta, rank: int, group, pad_size: int = 0) -> torch.Tensor:
    return _UndispatchFunc.apply(x_local, meta, rank, group, pad_size)
