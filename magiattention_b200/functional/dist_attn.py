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

Communication and compute overlap stage by stage through CUDA events.  Default engine on one NVSwitch domain
(comm/nvl_exchange.py): every transfer is a copy-engine pull over IPC-mapped peer memory on side streams, so no
communication kernel takes an SM from the attention grid (``sm_margin`` is therefore unnecessary); ranks are
ordered by epoch flags, not rendezvous.  Portable alternative (``MAGI_B200_KV_EXCHANGE=nccl``): grouped NCCL P2P
on NCCL's high-priority stream.  In both, stage i+1 transfers hide behind stage i compute and only the last
stage's dKV fold is exposed by construction (as in the reference, docs/source/blog/magi_attn.md:273-275).
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


class _Ready:
    """A finished 'work' around prefetched stage buffers (compute-only measurements)."""

    def __init__(self, bufs):
        self.bufs = bufs

    def wait_post_process(self, *a, **k):
        return self.bufs


class _SendOnly:
    """Stage in which this rank only sends: the transfer must be waited for, there is nothing to compute on."""

    send_only = True

    def __init__(self, work):
        self.work = work

    def wait_post_process(self, *a, **k):
        self.work.wait_post_process()
        return None


class DistAttnRuntime:
    """Holds the per-key metadata and implements the stage primitives (reference :141-239).

    Two exchange engines sit behind the same stage API:
      * ``nvl``  (default on one NVSwitch domain): copy-engine pulls over IPC-mapped peer memory, comm/nvl_exchange.py;
      * ``nccl`` (``MAGI_B200_KV_EXCHANGE=nccl``, or automatically when peer access is unavailable): group_cast /
        group_reduce as grouped ncclSend/ncclRecv, comm/group_collective.py."""

    def __init__(self, comm_meta: CommMeta, calc_meta: CalcMeta, cp_group: dist.ProcessGroup | None,
                 sm_margin: int = 0, high_precision_reduce: bool = False, all_comm_metas: list[CommMeta] | None = None,
                 deterministic: bool = False):
        self.comm_meta = comm_meta
        self.calc_meta = calc_meta
        self.cp_group = cp_group
        self.overlap_degree = calc_meta.overlap_degree
        if sm_margin != 0:
            raise NotImplementedError("sm_margin: the B200 exchange runs on copy engines and needs no SM reservation; the "
                                      "attention kernels are one-CTA-per-tile grids and cannot leave SMs idle")
        if deterministic:
            from .flex_flash_attn import _reject_deterministic
            _reject_deterministic()
        self.sm_margin = 0
        self.high_precision_reduce = high_precision_reduce
        self.cp_size = 1 if cp_group is None else dist.get_world_size(cp_group)
        # communication volume bookkeeping (rows of K == rows of V) for exposed-comm accounting
        self.remote_rows_per_stage = list(comm_meta.num_remote_kv_tokens_per_stage)
        self.all_comm_metas = all_comm_metas
        self.nvl = None
        self._stub: list | None = None  # compute-only mode (bench exposed-comm A/B): prefetched stage buffers, no transfers
        if self.cp_size > 1 and all_comm_metas is not None and cp_group is not None:
            from ..comm.nvl_exchange import NvlKVExchange, nvl_available

            if nvl_available(cp_group):  # collective: every rank builds its runtime for a key at the same point
                self.nvl = NvlKVExchange(cp_group, all_comm_metas, self.overlap_degree, calc_meta.seqlen_k_local)

    @property
    def engine(self) -> str:
        return "nvl" if self.nvl is not None else "nccl"

    def _stage_idle(self, st: int) -> bool:
        """True when this rank neither receives nor sends anything in stage ``st`` (decided from the transfer tables,
        never from a fetch handle being None: an owner that consumes nothing still has to take part in the reduce)."""
        arg = self.comm_meta.kv_group_collective_args_list[st]
        return self.remote_rows_per_stage[st] == 0 and sum(map(len, arg.dst_indices_list)) == 0

    # ---- communication ---------------------------------------------------------------------------
    def fetch_remote_kv(self, k: torch.Tensor, v: torch.Tensor, backward: bool = False) -> list[WorkWithPostProcessFn | None]:
        """Start the exchange of one pass: one async work per stage, all issued now
        (reference: _fetch_remote_kv :1463-1567).  A stage yields None when this rank receives no rows in it."""
        if self._stub is not None:
            return [_Ready(b) if b is not None else None for b in self._stub]
        if self.nvl is not None:
            return self.nvl.fetch(k, v, wire_esize=4 if self.high_precision_reduce else k.element_size(), backward=backward)
        works: list[WorkWithPostProcessFn | None] = []
        for st in range(self.overlap_degree):
            if self._stage_idle(st):
                works.append(None)
                continue
            arg = self.comm_meta.kv_group_collective_args_list[st]
            w = group_cast((k, v), None, group=self.cp_group, async_op=True, **arg.to_group_cast_args())
            if self.remote_rows_per_stage[st] == 0:  # send-only stage: keep the transfer alive, nothing to compute on
                w = _SendOnly(w)
            works.append(w)
        return works

    def mark_local_backward_done(self) -> None:
        if self._stub is None and self.nvl is not None:
            self.nvl.mark_local_done()

    def begin_reduce(self, dk: torch.Tensor, dv: torch.Tensor, wire: torch.dtype) -> None:
        """Owner side of the dKV return path (nvl engine): the pulls + folds of every stage on the reduce stream, ordered
        behind the local backward only.  Enqueued AFTER the stage loop (every partial-ready signal of this rank is in its
        stream by then — see NvlKVExchange.begin_reduce).  The NCCL engine reduces stage by stage instead."""
        if self._stub is None and self.nvl is not None:
            self.nvl.begin_reduce(dk, dv, wire)

    def partial_buffers(self, st: int, rk: torch.Tensor, rv: torch.Tensor, wire: torch.dtype):
        """Where the stage-``st`` backward kernel writes its partial dK,dV (wire dtype)."""
        if self._stub is None and self.nvl is not None:
            return self.nvl.partial_buffers(st, wire)
        return torch.empty(rk.shape, dtype=wire, device=rk.device), torch.empty(rv.shape, dtype=wire, device=rv.device)

    def reduce_partial_dkv(self, st: int, dk_part: torch.Tensor | None, dv_part: torch.Tensor | None, dk: torch.Tensor,
                           dv: torch.Tensor, wire: torch.dtype | None = None) -> WorkWithPostProcessFn | None:
        """Return a stage's partial dK,dV to their owners, summed into the owners' fp32 dk/dv
        (reference: _reduce_partial_dkv :2068-2195).  nvl: signal the owners (they pull); nccl: async group_reduce."""
        if self._stub is not None:
            return None
        if self.nvl is not None:
            if dk_part is not None:
                self.nvl.partial_ready(st)
            return None
        if self._stage_idle(st):
            return None
        arg = self.comm_meta.kv_group_collective_args_list[st]
        if dk_part is None:  # receive-only in the reduce (this rank consumed nothing of the stage): zero-row partials
            wire = wire or dk.dtype  # the receive staging takes the partials' dtype: it must be the wire dtype
            dk_part = torch.empty((0, *dk.shape[1:]), dtype=wire, device=dk.device)
            dv_part = torch.empty((0, *dv.shape[1:]), dtype=wire, device=dv.device)
        return group_reduce((dk_part, dv_part), (dk, dv), group=self.cp_group, async_op=True, reduce_op="sum",
                            acc_reduce=True, **arg.to_group_reduce_args())

    def finish_reduce(self, reduces: list) -> None:
        for rw in reduces:
            if rw is not None:
                rw.wait_post_process()
        if self._stub is None and self.nvl is not None:
            self.nvl.finish_reduce()

    def compute_only(self, stage_bufs: list | None):
        """Context manager for measurements: run the same kernels on prefetched stage buffers with every transfer and
        reduce stubbed out (the comm-disabled arm of the exposed-communication A/B)."""
        rt = self

        class _Ctx:
            def __enter__(self_):
                rt._stub = stage_bufs
            def __exit__(self_, *a):
                rt._stub = None
        return _Ctx()

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

    def attn_bwd(self, do, q, k, v, out, lse, arg: AttnArg, dq_acc, dk, dv, softmax_scale, softcap, sink=None, dsink=None,
                 row_stats=None):
        a = self._ffa_args(arg, q.device)
        _flex_flash_attn_backward(do, q, k, v, sink, "sh", out, lse, dq_acc, dk, dv, dsink, a["q_ranges"], a["k_ranges"],
                                  a["attn_type_map"], softmax_scale, softcap, sm_margin=self.sm_margin, row_stats=row_stats)


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
            bufs = w.wait_post_process()
            if bufs is None:
                continue
            rk, rv = bufs
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
        works = runtime.fetch_remote_kv(k, v, backward=True)
        dq = torch.zeros(q.shape, dtype=torch.float32, device=q.device)
        dk = torch.zeros(k.shape, dtype=torch.float32, device=k.device)
        dv = torch.zeros(v.shape, dtype=torch.float32, device=v.device)
        dsink = torch.zeros_like(sink, dtype=torch.float32) if sink is not None else None
        # every stage call of this pass sees the same (out, dout, lse): their row statistics are computed once
        row_stats = _lib.BwdRowStats()
        if not cm.local_attn_arg.can_skip or sink is not None:
            # the host stage also produces this rank's partial dsink (rows it owns; reference :663-713)
            runtime.attn_bwd(dout, q, k, v, out, lse, cm.local_attn_arg, dq, dk, dv, scale, softcap, sink=sink, dsink=dsink,
                             row_stats=row_stats)
        wire = torch.float32 if runtime.high_precision_reduce else k.dtype
        runtime.mark_local_backward_done()
        reduces: list = []
        for st in range(runtime.overlap_degree):
            w = works[st]
            bufs = w.wait_post_process() if w is not None else None
            if bufs is None:
                # nothing to compute on; the reduce may still need this rank as an owner (decided from the tables)
                reduces.append(runtime.reduce_partial_dkv(st, None, None, dk, dv, wire))
                continue
            rk, rv = bufs
            arg = cm.remote_attn_args_list[st]
            dk_p, dv_p = runtime.partial_buffers(st, rk, rv, wire)
            if arg.can_skip:
                dk_p.zero_(); dv_p.zero_()
            else:
                a = arg.to_ffa_args(q.device)
                # fresh (non-accumulating) partial dk/dv in the wire dtype; dq accumulates
                _lib.ffa_bwd(dout, q, rk, rv, out, lse, dq, dk_p, dv_p, a["q_ranges"], a["k_ranges"], a["attn_type_map"],
                             scale, softcap, accumulate_dkv=False, sm_margin=runtime.sm_margin, row_stats=row_stats)
            reduces.append(runtime.reduce_partial_dkv(st, dk_p, dv_p, dk, dv, wire))
        runtime.begin_reduce(dk, dv, wire)
        runtime.finish_reduce(reduces)
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
