"""Distributed (context-parallel) attention: the multi-stage overlap loop.

Mirror of the reference's L2 runtime for the hot path (magi_attention/functional/dist_attn.py:
DistAttnRuntime :141-2991, DistAttnFunc :2993-3605, dist_attn_func :3608-3680), B200-native:

  forward   issue group_cast of every stage's remote K,V up front (P2P batches on NCCL's stream, in stage
            order) -> local attention -> for each stage: wait its KV, run the kernel in ACCUMULATE mode so
            the stage merge happens in the kernel epilogue (no correct_attn_out_lse pass, no bf16 rounding
            of partial O between stages: fp32 (out, lse) accumulators) -> cast.
  backward  re-fetch KV the same way (nothing remote is saved) -> local backward (dq fp32 accumulator,
            local dk/dv written fp32) -> per stage: wait KV, backward against the stage buffer with the
            GLOBAL (out, lse) so partial grads are exact summands, then group_reduce the stage's partial
            dK,dV (16-bit wire dtype) back to their owners, asynchronously -> wait reduces -> cast.

Communication and compute overlap because the compute kernels are non-persistent (one CTA per SM, short
lived) and NCCL's stream is created high-priority (TORCH_NCCL_HIGH_PRIORITY=1, set by the runtime manager
unless overridden): P2P CTAs are scheduled as attention CTAs retire, stage i+1 transfers hide behind stage
i compute, and only the last stage's dKV reduce is exposed by construction (as in the reference,
docs/source/blog/magi_attn.md:273-275).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from .. import _lib
from ..comm.group_collective import group_cast, group_reduce
from ..comm.work import WorkWithPostProcessFn
from ..common.forward_meta import AttnForwardMeta
from ..meta.planner import AttnArg, CalcMeta, CommMeta
from .flex_flash_attn import _flex_flash_attn_backward, _flex_flash_attn_forward

__all__ = ["DistAttnRuntime", "DistAttnFunc", "dist_attn_func"]


class DistAttnRuntime:
    """Holds the per-key metadata and implements the stage primitives (reference :141-239)."""

    def __init__(self, comm_meta: CommMeta, calc_meta: CalcMeta, cp_group: dist.ProcessGroup | None,
                 sm_margin: int = 0, high_precision_reduce: bool = False, all_comm_metas: list[CommMeta] | None = None):
        self.comm_meta = comm_meta
        self.calc_meta = calc_meta
        self.cp_group = cp_group
        self.overlap_degree = calc_meta.overlap_degree
        self.sm_margin = sm_margin
        self.high_precision_reduce = high_precision_reduce
        self.cp_size = 1 if cp_group is None else dist.get_world_size(cp_group)
        # communication volume bookkeeping (rows of K == rows of V) for exposed-comm accounting
        self.remote_rows_per_stage = list(comm_meta.num_remote_kv_tokens_per_stage)
        # opt-in experimental engine: KV pulled over NVLink peer memory instead of NCCL P2P (comm/peer_exchange.py)
        self.all_comm_metas = all_comm_metas
        self._peer_fetchers: dict = {}

    # ---- communication ---------------------------------------------------------------------------
    def fetch_remote_kv(self, k: torch.Tensor, v: torch.Tensor) -> list[WorkWithPostProcessFn | None]:
        """One async group_cast per stage, all issued now (reference: _fetch_remote_kv :1463-1567)."""
        if os.environ.get("MAGI_B200_KV_EXCHANGE", "nccl") == "peer" and self.cp_size > 1 and self.all_comm_metas is not None:
            return self._peer_fetcher(k).fetch(k, v)
        works: list[WorkWithPostProcessFn | None] = []
        for st in range(self.overlap_degree):
            if self.calc_meta.remote_attn_args_list[st].can_skip and self.remote_rows_per_stage[st] == 0 \
                    and sum(map(len, self.comm_meta.kv_group_collective_args_list[st].dst_indices_list)) == 0:
                works.append(None)
                continue
            arg = self.comm_meta.kv_group_collective_args_list[st]
            works.append(group_cast((k, v), None, group=self.cp_group, async_op=True, **arg.to_group_cast_args()))
        return works

    def _peer_fetcher(self, k: torch.Tensor):
        from ..comm.peer_exchange import PeerKVFetcher

        key = (k.dtype, k.shape[1], k.shape[2], k.device)
        f = self._peer_fetchers.get(key)
        if f is None:  # collective: every rank builds it on its first fetch with this geometry
            per_stage = [[cm.kv_group_collective_args_list[st] for cm in self.all_comm_metas] for st in range(self.overlap_degree)]
            f = PeerKVFetcher(self.cp_group, dist.get_rank(self.cp_group), per_stage, k.shape[0], k.shape[1], k.shape[2], k.dtype,
                              k.device)
            self._peer_fetchers[key] = f
        return f

    def reduce_partial_dkv(self, st: int, dk_part: torch.Tensor, dv_part: torch.Tensor, dk: torch.Tensor,
                           dv: torch.Tensor) -> WorkWithPostProcessFn:
        """Async group_reduce(sum) of a stage's partial dK,dV into the owners' fp32 dk/dv
        (reference: _reduce_partial_dkv :2068-2195)."""
        arg = self.comm_meta.kv_group_collective_args_list[st]
        return group_reduce((dk_part, dv_part), (dk, dv), group=self.cp_group, async_op=True, reduce_op="sum",
                            acc_reduce=True, **arg.to_group_reduce_args())

    def reduce_max_logits(self, max_logits: torch.Tensor | None) -> torch.Tensor | None:
        """Element-wise MAX over the cp group (reference: reduce_max_logits :532-555)."""
        if max_logits is not None and self.cp_group is not None and self.cp_size > 1:
            dist.all_reduce(max_logits, op=dist.ReduceOp.MAX, group=self.cp_group)
        return max_logits

    def reduce_partial_dsink(self, dsink: torch.Tensor | None) -> torch.Tensor | None:
        """The sink is replicated, every rank holds the partial gradient of its own rows.  Like the reference
        (env/comm.py:137-168, dist_attn.py:882-896) the cross-rank reduction is opt-in through
        MAGI_ATTENTION_DSINK_ALL_REDUCE_OP in {none (default: the training framework reduces parameters), sum, avg}."""
        op = os.environ.get("MAGI_ATTENTION_DSINK_ALL_REDUCE_OP", "none")
        assert op in ("none", "sum", "avg"), f"Invalid value of MAGI_ATTENTION_DSINK_ALL_REDUCE_OP: {op}"
        if dsink is not None and op != "none" and self.cp_group is not None and self.cp_size > 1:
            dist.all_reduce(dsink, op=dist.ReduceOp.SUM, group=self.cp_group)
            if op == "avg":
                dsink /= self.cp_size
        return dsink

    # ---- compute -----------------------------------------------------------------------------------
    @staticmethod
    def _ffa_args(arg: AttnArg, device) -> dict:
        if arg.can_skip:  # no slices: the kernels still define uncovered rows (out 0 / lse -inf) and apply the sink
            e = torch.empty((0, 2), dtype=torch.int32, device=device)
            return {"q_ranges": e, "k_ranges": e, "attn_type_map": torch.empty((0,), dtype=torch.int32, device=device)}
        return arg.to_ffa_args(device)

    def attn_fwd(self, q, k, v, arg: AttnArg, out_acc, lse_acc, softmax_scale, softcap, sink=None, max_logits=None):
        """One stage into the fp32 accumulators.  ``sink`` is passed on the host stage only: it must be applied
        exactly once (reference :1278-1280), and then also when the host stage has no slices (:2380-2408)."""
        if arg.can_skip and sink is None:
            return
        a = self._ffa_args(arg, q.device)
        _flex_flash_attn_forward(q, k, v, sink, "sh", out_acc, lse_acc, a["q_ranges"], a["k_ranges"], a["attn_type_map"],
                                 softmax_scale, softcap, torch.float32, sm_margin=self.sm_margin,
                                 return_max_logits=max_logits is not None, max_logits=max_logits)

    def attn_bwd(self, do, q, k, v, out, lse, arg: AttnArg, dq_acc, dk, dv, softmax_scale, softcap, sink=None, dsink=None):
        a = self._ffa_args(arg, q.device)
        _flex_flash_attn_backward(do, q, k, v, sink, "sh", out, lse, dq_acc, dk, dv, dsink, a["q_ranges"], a["k_ranges"],
                                  a["attn_type_map"], softmax_scale, softcap, sm_margin=self.sm_margin)


class DistAttnFunc(torch.autograd.Function):
    """Reference: DistAttnFunc.forward :2997-3194 / backward :3455-3605."""

    @staticmethod
    def forward(ctx, q, k, v, sink, runtime: DistAttnRuntime, softmax_scale, softcap, return_max_logits):
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        cm = runtime.calc_meta
        works = runtime.fetch_remote_kv(k, v)
        out_acc = torch.zeros(q.shape, dtype=torch.float32, device=q.device)
        lse_acc = torch.full((q.shape[0], q.shape[1]), float("-inf"), dtype=torch.float32, device=q.device)
        max_logits = (torch.full((q.shape[1],), float("-inf"), dtype=torch.float32, device=q.device)
                      if return_max_logits else None)
        runtime.attn_fwd(q, k, v, cm.local_attn_arg, out_acc, lse_acc, softmax_scale, softcap, sink=sink, max_logits=max_logits)
        for st in range(runtime.overlap_degree):
            w = works[st]
            if w is None:
                continue
            rk, rv = w.wait_post_process()
            runtime.attn_fwd(q, rk, rv, cm.remote_attn_args_list[st], out_acc, lse_acc, softmax_scale, softcap,
                             max_logits=max_logits)
        max_logits = runtime.reduce_max_logits(max_logits)
        out = out_acc.to(q.dtype)
        ctx.save_for_backward(q, k, v, sink, out, lse_acc)
        ctx.runtime, ctx.softmax_scale, ctx.softcap = runtime, softmax_scale, softcap
        ctx.mark_non_differentiable(lse_acc)
        if max_logits is not None:
            ctx.mark_non_differentiable(max_logits)
        return out, lse_acc, max_logits

    @staticmethod
    def backward(ctx, dout, *args):
        q, k, v, sink, out, lse = ctx.saved_tensors
        runtime: DistAttnRuntime = ctx.runtime
        cm = runtime.calc_meta
        scale, softcap = ctx.softmax_scale, ctx.softcap
        dout = dout.contiguous()
        works = runtime.fetch_remote_kv(k, v)
        dq = torch.zeros(q.shape, dtype=torch.float32, device=q.device)
        dk = torch.zeros(k.shape, dtype=torch.float32, device=k.device)
        dv = torch.zeros(v.shape, dtype=torch.float32, device=v.device)
        dsink = torch.zeros_like(sink, dtype=torch.float32) if sink is not None else None
        if not cm.local_attn_arg.can_skip or sink is not None:
            # the host stage also produces this rank's partial dsink (rows it owns; reference :663-713)
            runtime.attn_bwd(dout, q, k, v, out, lse, cm.local_attn_arg, dq, dk, dv, scale, softcap, sink=sink, dsink=dsink)
        reduces: list[WorkWithPostProcessFn] = []
        wire = torch.float32 if runtime.high_precision_reduce else k.dtype
        for st in range(runtime.overlap_degree):
            w = works[st]
            if w is None:
                continue
            rk, rv = w.wait_post_process()
            arg = cm.remote_attn_args_list[st]
            if arg.can_skip:
                dk_p = torch.zeros(rk.shape, dtype=wire, device=rk.device)
                dv_p = torch.zeros(rv.shape, dtype=wire, device=rv.device)
            else:
                dk_p = torch.empty(rk.shape, dtype=wire, device=rk.device)
                dv_p = torch.empty(rv.shape, dtype=wire, device=rv.device)
                a = arg.to_ffa_args(q.device)
                # fresh (non-accumulating) partial dk/dv in the wire dtype; dq accumulates
                _lib.ffa_bwd(dout, q, rk, rv, out, lse, dq, dk_p, dv_p, a["q_ranges"], a["k_ranges"], a["attn_type_map"],
                             scale, softcap, accumulate_dkv=False, sm_margin=runtime.sm_margin)
            reduces.append(runtime.reduce_partial_dkv(st, dk_p, dv_p, dk, dv))
        for rw in reduces:
            rw.wait_post_process()
        dsink = runtime.reduce_partial_dsink(dsink)
        if dsink is not None:
            dsink = dsink.to(sink.dtype)
        return dq.to(q.dtype), dk.to(k.dtype), dv.to(v.dtype), dsink, None, None, None, None


def dist_attn_func(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, dist_attn_runtime: DistAttnRuntime,
                   sink: torch.Tensor | None = None, softmax_scale: float | None = None, softcap: float = 0.0,
                   return_max_logits: bool = False) -> tuple[torch.Tensor, AttnForwardMeta]:
    """Local shards in, local attention out (reference signature: dist_attn.py:3608-3680)."""
    if softcap != 0.0:
        raise RuntimeError("B200 backend does not support softcap")
    if sink is not None and sink.dim() != 2:
        raise RuntimeError('dist_attn_func takes a replicated sink with layout "sh": [num_tokens_sink, num_heads_q]')
    if softmax_scale is None:
        softmax_scale = q.shape[-1] ** (-0.5)
    out, lse, max_logits = DistAttnFunc.apply(q, k, v, sink, dist_attn_runtime, softmax_scale, softcap, return_max_logits)
    return out, AttnForwardMeta(lse=lse, max_logits=max_logits)
