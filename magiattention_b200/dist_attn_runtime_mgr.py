"""Runtime key / manager: the L4 entry points the API layer (and Megatron / torch_native glue) call.

Same names, parameter lists, order and defaults as the reference (magi_attention/dist_attn_runtime_mgr.py):
``DistAttnRuntimeKey`` :61-120, ``DistAttnRuntimeMgr`` :122-420 (``dispatch_qo`` / ``dispatch_kv`` / ``undispatch_qo`` /
``undispatch_kv`` :162-199, ``calc_attn`` :246-265, ``get_xattn_args`` :267-355, ``get_position_ids(attn_role)`` :357-377),
``init_dist_attn_runtime_key`` :484-530, ``init_dist_attn_runtime_mgr`` :545-1010.  The signatures are pinned against
the reference module by tests/test_api_signatures.py.

A key identifies (padded mask, shapes, cp group, config, behaviour flags); the manager owns the plan of THIS rank
(dispatch meta, per-stage kernel slices, transfer tables) and the DistAttnRuntime that executes it.  As in the reference the
q/k ranges and ``total_seqlen_*`` handed to these functions already include the padding block (api/magi_attn_interface.py
:664-676 applies it before building the key).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Any

import torch
import torch.distributed as dist

from .common.enum import AttnMaskType, AttnRole
from .common.forward_meta import AttnForwardMeta
from .common.ranges import AttnRanges
from .config import DistAttnConfig
from .functional.dist_attn import DistAttnRuntime, dist_attn_func
from .functional.dispatch import dispatch_func, undispatch_func
from .functional.roll import roll_func, roll_simple_func
from .meta.planner import DispatchMeta, make_attn_meta, make_dispatch_meta

__all__ = ["DistAttnRuntimeKey", "DistAttnRuntimeMgr", "init_dist_attn_runtime_key", "init_dist_attn_runtime_mgr",
           "check_flag_comb"]

KERNEL_BACKEND = "b200_ffa"


def _env_flag(name: str) -> bool:
    return os.environ.get(name, "0") == "1"


@dataclass(frozen=True)
class DistAttnRuntimeKey:
    """Field for field the reference's key (:61-84); hashable, so equal masks share one planned runtime."""

    q_ranges: AttnRanges
    k_ranges: AttnRanges
    attn_mask_type: tuple[AttnMaskType, ...]
    total_seqlen_q: int
    total_seqlen_k: int
    num_heads_q: int
    num_heads_kv: int
    head_dim: int
    pad_size: int
    chunk_size: int
    cp_group: Any  # dist.ProcessGroup | None (None: single-process cp=1, an extension for tests / single-GPU use)
    cp_mesh: Any   # DeviceMesh | None
    dist_attn_config: DistAttnConfig
    uneven_shard: bool

    # flags that might influence the runtime behavior
    is_deterministic_mode_enable: bool
    is_hierarchical_comm_enable: bool
    is_qo_comm_enable: bool
    is_native_grpcoll_enable: bool
    is_flatten_head_groups_enable: bool
    kernel_backend: str
    precision: Any
    is_auto_range_merge_enable: bool

    # (q_ranges, k_ranges, types) of the mask whose dispatch solution this key reuses; None = its own mask
    # (reference: make_*_key_for_new_mask_after_dispatch pass ref_dispatch_meta_q/k, api/magi_attn_interface.py:1476-1530)
    dispatch_mask: tuple | None = None

    # -- conveniences used inside this package ------------------------------------------------------------------
    @property
    def cp_size(self) -> int:
        return 1 if self.cp_group is None else dist.get_world_size(self.cp_group)

    @property
    def config(self) -> DistAttnConfig:
        return self.dist_attn_config

    @property
    def int_types(self) -> tuple[int, ...]:
        return tuple(t.to_int_type() for t in self.attn_mask_type)


def check_flag_comb() -> None:
    """Reject the behaviour switches of the reference that this backend does not implement (reference check_flag_comb
    :425-482 validates ITS combinations; every one of these selects a subsystem that SURVEY §2 places out of scope)."""
    for env_name, what in (("MAGI_ATTENTION_HIERARCHICAL_COMM", "hierarchical comm"),
                           ("MAGI_ATTENTION_QO_COMM", "q/o communication (dynamic solver)"),
                           ("MAGI_ATTENTION_NATIVE_GRPCOLL", "the NVSHMEM native group collective"),
                           ("MAGI_ATTENTION_FLATTEN_HEAD_GROUPS", "flattened head groups")):
        if _env_flag(env_name):
            raise NotImplementedError(f"{env_name}=1: {what} is not part of the B200 backend")
    if _env_flag("MAGI_ATTENTION_DETERMINISTIC_MODE"):
        from .functional.flex_flash_attn import DETERMINISTIC_SUPPORTED

        if not DETERMINISTIC_SUPPORTED:
            raise NotImplementedError("MAGI_ATTENTION_DETERMINISTIC_MODE=1 is not supported by this build")


def _as_mask_types(attn_mask_type, n: int) -> tuple[AttnMaskType, ...]:
    if not isinstance(attn_mask_type, (list, tuple)):
        attn_mask_type = [attn_mask_type] * n
    out = []
    for t in attn_mask_type:
        if isinstance(t, AttnMaskType):
            out.append(t)
        elif isinstance(t, str):
            out.append(AttnMaskType(t))
        else:
            out.append(AttnMaskType.from_int_type(int(t)))
    return tuple(out)


def init_dist_attn_runtime_key(
    q_ranges: AttnRanges,
    k_ranges: AttnRanges,
    attn_mask_type: list[AttnMaskType],
    total_seqlen_q: int,
    total_seqlen_k: int,
    num_heads_q: int,
    num_heads_kv: int,
    head_dim: int,
    pad_size: int,
    chunk_size: int,
    cp_group: dist.ProcessGroup,
    cp_mesh,
    dist_attn_config: DistAttnConfig,
) -> DistAttnRuntimeKey:
    """Initialize DistAttnRuntimeKey (reference :484-530).  A key that rides on another mask's dispatch solution is derived
    from the result with ``dataclasses.replace(key, dispatch_mask=...)`` (api.make_*_key_for_new_mask_after_dispatch)."""
    check_flag_comb()
    types = _as_mask_types(attn_mask_type, len(q_ranges))
    assert len(types) == len(q_ranges) == len(k_ranges)
    return DistAttnRuntimeKey(
        q_ranges=q_ranges, k_ranges=k_ranges, attn_mask_type=types, total_seqlen_q=total_seqlen_q,
        total_seqlen_k=total_seqlen_k, num_heads_q=num_heads_q, num_heads_kv=num_heads_kv, head_dim=head_dim,
        pad_size=pad_size, chunk_size=chunk_size, cp_group=cp_group, cp_mesh=cp_mesh, dist_attn_config=dist_attn_config,
        uneven_shard=dist_attn_config.dispatch_config.uneven_shard,
        is_deterministic_mode_enable=_env_flag("MAGI_ATTENTION_DETERMINISTIC_MODE"),
        is_hierarchical_comm_enable=False, is_qo_comm_enable=False, is_native_grpcoll_enable=False,
        is_flatten_head_groups_enable=False, kernel_backend=KERNEL_BACKEND, precision=None,
        is_auto_range_merge_enable=_env_flag("MAGI_ATTENTION_AUTO_RANGE_MERGE"),
    )


class DistAttnRuntimeMgr:
    """Reference :122-420.  ``dispatch_meta_q`` and ``dispatch_meta_k`` are the same object for self-attention."""

    def __init__(
        self,
        dispatch_meta_q: DispatchMeta,
        dispatch_meta_k: DispatchMeta,
        dist_attn_config: DistAttnConfig,
        attn_solver,
        dist_attn_runtime: DistAttnRuntime,
        num_heads_q: int,
        num_heads_kv: int,
        head_dim: int,
        cp_group: dist.ProcessGroup,
        *,
        ref_q_ranges: AttnRanges,
        ref_k_ranges: AttnRanges,
        is_same_source: bool,
        is_q_permutable: bool,
        is_k_permutable: bool,
    ):
        self.dispatch_meta_q = dispatch_meta_q
        self.dispatch_meta_k = dispatch_meta_k
        self.dist_attn_config = dist_attn_config
        # the reference stores its DistAttnSolver here; this backend's planner is closed-form, its product is the pair
        # (comm metas, calc metas) of every rank
        self.attn_solver = attn_solver
        self.dist_attn_runtime = dist_attn_runtime
        self.num_heads_q, self.num_heads_kv, self.head_dim = num_heads_q, num_heads_kv, head_dim
        self.cp_group = cp_group
        self.cp_size = 1 if cp_group is None else dist.get_world_size(cp_group)
        self.cp_rank = 0 if cp_group is None else dist.get_rank(cp_group)
        self.ref_q_ranges = ref_q_ranges
        self.ref_k_ranges = ref_k_ranges
        self.is_same_source = is_same_source
        self.is_q_permutable = is_q_permutable
        self.is_k_permutable = is_k_permutable
        self._q_position_ids: torch.Tensor | None = None
        self._k_position_ids: torch.Tensor | None = None
        # filled by init_dist_attn_runtime_mgr: the plan of every rank (FLOP / balance accounting, NVLink pull tables)
        self.comm_meta = dist_attn_runtime.comm_meta
        self.calc_meta = dist_attn_runtime.calc_meta
        self.all_calc_metas: list = []
        self.all_comm_metas: list = []

    @property
    def dispatch_meta(self) -> DispatchMeta:
        return self.dispatch_meta_q

    # ---- dispatch / undispatch (reference :162-199): tensors already padded to the key's total_seqlen ---------------
    def dispatch_qo(self, q_or_o: torch.Tensor) -> torch.Tensor:
        return dispatch_func(q_or_o, self.dispatch_meta_q, self.cp_rank, self.cp_group)

    def dispatch_kv(self, k_or_v: torch.Tensor) -> torch.Tensor:
        return dispatch_func(k_or_v, self.dispatch_meta_k, self.cp_rank, self.cp_group)

    def undispatch_qo(self, q_or_o: torch.Tensor, is_partial_grad: bool = False) -> torch.Tensor:
        return undispatch_func(q_or_o, self.dispatch_meta_q, self.cp_rank, self.cp_group, is_partial_grad=is_partial_grad)

    def undispatch_kv(self, k_or_v: torch.Tensor, is_partial_grad: bool = False) -> torch.Tensor:
        return undispatch_func(k_or_v, self.dispatch_meta_k, self.cp_rank, self.cp_group, is_partial_grad=is_partial_grad)

    def roll(self, x: torch.Tensor, shift: int, dim: int) -> torch.Tensor:
        """Cyclic roll of a dispatched tensor along the global sequence (reference :200-221)."""
        return roll_func(x, shift, self.dispatch_meta_q, self.cp_rank, self.cp_group, dim)

    def roll_simple(self, x: torch.Tensor, shift: int, dim: int) -> torch.Tensor:
        return roll_simple_func(x, shift, self.dispatch_meta_q, self.cp_rank, self.cp_group, dim)

    # ---- attention (reference :246-265) --------------------------------------------------------------
    def calc_attn(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, sink: torch.Tensor | None = None,
                  softmax_scale: float | None = None, softcap: float = 0.0,
                  return_max_logits: bool = False) -> tuple[torch.Tensor, AttnForwardMeta]:
        return dist_attn_func(q, k, v, self.dist_attn_runtime, sink=sink, softmax_scale=softmax_scale, softcap=softcap,
                              return_max_logits=return_max_logits)

    def get_xattn_args(self, xattn_k_ranges: AttnRanges, attn_mask_type, return_host_only: bool = False):
        """Cross-attention arguments for an external, un-dispatched key source (reference :267-355): needs
        ``is_same_source=False`` runtimes, which this backend does not plan (cross-attn out of scope)."""
        raise NotImplementedError("cross-attention (get_xattn_args) is not part of the B200 backend")

    def get_position_ids(self, attn_role: AttnRole = AttnRole.QUERY) -> torch.Tensor:
        """Global position of every local token w.r.t. ``attn_role`` (reference :357-377)."""
        if attn_role == AttnRole.QUERY:
            if self._q_position_ids is None:
                self._q_position_ids = self.dispatch_meta_q.position_ids(self.cp_rank)
            return self._q_position_ids
        if attn_role in (AttnRole.KEY, AttnRole.VALUE):
            if self._k_position_ids is None:
                self._k_position_ids = self.dispatch_meta_k.position_ids(self.cp_rank)
            return self._k_position_ids
        raise ValueError(f"Invalid attn role: {attn_role}")


def init_dist_attn_runtime_mgr(
    q_ranges: AttnRanges,
    k_ranges: AttnRanges,
    attn_mask_type: list[AttnMaskType],
    total_seqlen_q: int,
    total_seqlen_k: int,
    num_heads_q: int,
    num_heads_kv: int,
    head_dim: int,
    chunk_size: int,
    cp_group: dist.ProcessGroup,
    cp_mesh=None,
    dist_attn_config: DistAttnConfig = DistAttnConfig(),
    is_same_source: bool = True,
    is_q_permutable: bool = True,
    is_k_permutable: bool = True,
    ref_dispatch_meta_q: DispatchMeta | None = None,
    ref_dispatch_meta_k: DispatchMeta | None = None,
) -> DistAttnRuntimeMgr:
    """Plan the mask for ``cp_group`` and build the runtime (reference :545-1010).

    ``q_ranges`` / ``k_ranges`` / ``total_seqlen_*`` include the padding block.  ``ref_dispatch_meta_q/k``: reuse an existing
    dispatch solution (a second mask on already-dispatched tensors) instead of balancing this mask."""
    check_flag_comb()
    if not (is_same_source and is_q_permutable and is_k_permutable):
        raise NotImplementedError("only self-attention (is_same_source, q and k permutable) is planned by the B200 backend")
    if cp_mesh is not None and getattr(cp_mesh, "ndim", 1) != 1:
        raise NotImplementedError("2-D cp_mesh (hierarchical comm) is not supported: one NVSwitch domain is flat")
    assert total_seqlen_q == total_seqlen_k, "self-attention: total_seqlen_q must equal total_seqlen_k"
    if (ref_dispatch_meta_q is None) != (ref_dispatch_meta_k is None):
        raise ValueError("ref_dispatch_meta_q and ref_dispatch_meta_k must be given together")
    cp_size = 1 if cp_group is None else dist.get_world_size(cp_group)
    cp_rank = 0 if cp_group is None else dist.get_rank(cp_group)
    types = [t.to_int_type() for t in _as_mask_types(attn_mask_type, len(q_ranges))]
    qn, kn = q_ranges.to_naive_ranges(), k_ranges.to_naive_ranges()
    slices = [(q[0], q[1], k[0], k[1], t) for q, k, t in zip(qn, kn, types) if q[1] > q[0] and k[1] > k[0]]
    if ref_dispatch_meta_q is not None:
        dm = ref_dispatch_meta_q
        assert dm.total_seqlen == total_seqlen_q and dm.cp_size == cp_size, "ref_dispatch_meta does not fit this mask"
    else:
        dm = make_dispatch_meta(slices, total_seqlen_q, chunk_size, cp_size, dist_attn_config.dispatch_config.planner_alg)
    degree = dist_attn_config.overlap_config.num_stages(cp_size)
    comm_metas, calc_metas = make_attn_meta(slices, dm, degree)
    # fp32 wire dtype for the partial dK/dV reduce: config field, or the reference's env switch (env/comm.py:107-121)
    hp_reduce = dist_attn_config.high_precision_reduce or _env_flag("MAGI_ATTENTION_BACKWARD_HIGH_PRECISION_REDUCE")
    runtime = DistAttnRuntime(comm_metas[cp_rank], calc_metas[cp_rank], cp_group, high_precision_reduce=hp_reduce,
                              all_comm_metas=comm_metas, deterministic=_env_flag("MAGI_ATTENTION_DETERMINISTIC_MODE"))
    mgr = DistAttnRuntimeMgr(dm, dm, dist_attn_config, (comm_metas, calc_metas), runtime, num_heads_q, num_heads_kv, head_dim,
                             cp_group, ref_q_ranges=q_ranges, ref_k_ranges=k_ranges, is_same_source=is_same_source,
                             is_q_permutable=is_q_permutable, is_k_permutable=is_k_permutable)
    mgr.all_calc_metas, mgr.all_comm_metas = calc_metas, comm_metas
    return mgr
