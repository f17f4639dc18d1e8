"""Runtime key / manager: the entry points the API layer calls
(reference: magi_attention/dist_attn_runtime_mgr.py — DistAttnRuntimeKey :61, DistAttnRuntimeMgr
:122-377, init_dist_attn_runtime_key :484, init_dist_attn_runtime_mgr :545).

A key identifies (mask, shapes, cp group, config); the manager owns the plan for THIS rank
(dispatch meta, per-stage kernel slices, transfer tables) and the DistAttnRuntime that executes it.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch
import torch.distributed as dist

from .common.enum import AttnMaskType
from .common.forward_meta import AttnForwardMeta
from .common.ranges import AttnRanges
from .config import DistAttnConfig
from .functional.dist_attn import DistAttnRuntime, dist_attn_func
from .functional.dispatch import dispatch_func, undispatch_func
from .functional.roll import roll_func, roll_simple_func
from .meta.planner import DispatchMeta, make_attn_meta, make_dispatch_meta

__all__ = ["DistAttnRuntimeKey", "DistAttnRuntimeMgr", "init_dist_attn_runtime_key", "init_dist_attn_runtime_mgr"]


@dataclass(frozen=True)
class DistAttnRuntimeKey:
    q_ranges: tuple[tuple[int, int], ...]
    k_ranges: tuple[tuple[int, int], ...]
    attn_mask_type: tuple[int, ...]
    total_seqlen_q: int
    total_seqlen_k: int
    pad_size: int
    chunk_size: int
    num_heads_q: int
    num_heads_kv: int
    head_dim: int
    cp_group_id: int
    cp_size: int
    config: DistAttnConfig
    # (q_ranges, k_ranges, types) of the mask whose dispatch solution this key reuses; None = its own mask
    # (reference: make_*_key_for_new_mask_after_dispatch pass ref_dispatch_meta_q/k, magi_attn_interface.py:1476-1530)
    dispatch_mask: tuple | None = None


def _mask_type_to_int(t) -> int:
    if isinstance(t, AttnMaskType):
        return t.to_int_type()
    if isinstance(t, str):
        return AttnMaskType(t).to_int_type()
    return int(t)


def init_dist_attn_runtime_key(q_ranges: AttnRanges, k_ranges: AttnRanges, attn_mask_type, total_seqlen_q: int,
                               total_seqlen_k: int, pad_size: int, chunk_size: int, cp_group: dist.ProcessGroup | None,
                               num_heads_q: int, num_heads_kv: int, head_dim: int,
                               dist_attn_config: DistAttnConfig = DistAttnConfig(),
                               dispatch_mask: tuple | None = None) -> DistAttnRuntimeKey:
    n = len(q_ranges)
    if not isinstance(attn_mask_type, (list, tuple)):
        attn_mask_type = [attn_mask_type] * n
    types = tuple(_mask_type_to_int(t) for t in attn_mask_type)
    assert len(types) == n == len(k_ranges)
    cp_size = 1 if cp_group is None else dist.get_world_size(cp_group)
    return DistAttnRuntimeKey(tuple(q_ranges.to_naive_ranges()), tuple(k_ranges.to_naive_ranges()), types,
                              total_seqlen_q, total_seqlen_k, pad_size, chunk_size, num_heads_q, num_heads_kv, head_dim,
                              id(cp_group) if cp_group is not None else 0, cp_size, dist_attn_config, dispatch_mask)


class DistAttnRuntimeMgr:
    def __init__(self, key: DistAttnRuntimeKey, cp_group: dist.ProcessGroup | None):
        assert key.total_seqlen_q == key.total_seqlen_k, "self-attention (same-source q/k) only"
        self.key = key
        self.cp_group = cp_group
        self.cp_size = key.cp_size
        self.cp_rank = 0 if cp_group is None else dist.get_rank(cp_group)
        total = key.total_seqlen_q + key.pad_size
        self.total_seqlen = total
        slices = [(q[0], q[1], k[0], k[1], t) for q, k, t in zip(key.q_ranges, key.k_ranges, key.attn_mask_type)
                  if q[1] > q[0] and k[1] > k[0]]
        dslices = slices
        if key.dispatch_mask is not None:  # a second mask riding on another mask's dispatch solution
            dq, dk, dt = key.dispatch_mask
            dslices = [(q[0], q[1], k[0], k[1], t) for q, k, t in zip(dq, dk, dt) if q[1] > q[0] and k[1] > k[0]]
        self.dispatch_meta: DispatchMeta = make_dispatch_meta(dslices, total, key.chunk_size, self.cp_size,
                                                              key.config.dispatch_config.planner_alg)
        degree = key.config.overlap_config.num_stages(self.cp_size)
        comm_metas, calc_metas = make_attn_meta(slices, self.dispatch_meta, degree)
        self.comm_meta, self.calc_meta = comm_metas[self.cp_rank], calc_metas[self.cp_rank]
        self.all_calc_metas = calc_metas  # kept for FLOP / balance accounting
        self.all_comm_metas = comm_metas
        # fp32 wire dtype for the partial dK/dV reduce: config field, or the reference's env switch (env/comm.py:107-121)
        hp_reduce = key.config.high_precision_reduce or os.environ.get("MAGI_ATTENTION_BACKWARD_HIGH_PRECISION_REDUCE", "0") == "1"
        self.dist_attn_runtime = DistAttnRuntime(self.comm_meta, self.calc_meta, cp_group, high_precision_reduce=hp_reduce,
                                                 all_comm_metas=comm_metas)
        self._position_ids: torch.Tensor | None = None

    # ---- dispatch / undispatch (reference :162-245) ------------------------------------------------
    def dispatch_qo(self, x: torch.Tensor) -> torch.Tensor:
        return dispatch_func(x, self.dispatch_meta, self.cp_rank, self.cp_group, self.key.pad_size)

    dispatch_kv = dispatch_qo

    def undispatch_qo(self, x: torch.Tensor) -> torch.Tensor:
        return undispatch_func(x, self.dispatch_meta, self.cp_rank, self.cp_group, self.key.pad_size)

    undispatch_kv = undispatch_qo

    def roll(self, x: torch.Tensor, shift: int, dim: int) -> torch.Tensor:
        """Cyclic roll of a dispatched tensor along the global sequence (reference :200-221)."""
        return roll_func(x, shift, self.dispatch_meta, self.cp_rank, self.cp_group, dim)

    def roll_simple(self, x: torch.Tensor, shift: int, dim: int) -> torch.Tensor:
        return roll_simple_func(x, shift, self.dispatch_meta, self.cp_rank, self.cp_group, dim)

    # ---- attention (reference :246-265) --------------------------------------------------------------
    def calc_attn(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, sink: torch.Tensor | None = None,
                  softmax_scale: float | None = None, softcap: float = 0.0,
                  return_max_logits: bool = False) -> tuple[torch.Tensor, AttnForwardMeta]:
        return dist_attn_func(q, k, v, self.dist_attn_runtime, sink=sink, softmax_scale=softmax_scale, softcap=softcap,
                              return_max_logits=return_max_logits)

    def get_position_ids(self, device=None) -> torch.Tensor:
        if self._position_ids is None:
            self._position_ids = self.dispatch_meta.position_ids(self.cp_rank)
        return self._position_ids if device is None else self._position_ids.to(device)


def init_dist_attn_runtime_mgr(key: DistAttnRuntimeKey, cp_group: dist.ProcessGroup | None) -> DistAttnRuntimeMgr:
    # NCCL's stream must be high priority so that P2P CTAs win SM slots from the attention grid
    os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
    return DistAttnRuntimeMgr(key, cp_group)
