"""DistAttnConfig — user-facing knobs, same classes and field names as the reference
(magi_attention/config.py:54-71; DispatchConfig + *DispatchAlg: meta/solver/dispatch_solver.py:32-367;
OverlapConfig + *OverlapAlg: meta/solver/overlap_solver.py:26-150).

Everything here is frozen and hashable: the config is part of the runtime key.

What the B200 planner does with the knobs:
  * dispatch algorithms — ``MinHeapDispatchAlg`` (default: chunks by decreasing mask area to the least-loaded rank, equal
    chunk counts), ``SequentialDispatchAlg`` (contiguous blocks of chunks per rank), ``SortedSequentialSelectAlg`` (area-sorted
    chunks dealt out in snake order).  ``LBDispatchAlg`` / ``DPDispatchAlg`` / ``BSDispatchAlg`` / ``ToppHeapDispatchAlg`` are
    other balancers for the same objective (minimise the max per-rank area): they are accepted and served by the min-heap
    balancer — the attention result never depends on the dispatch solution, only the load balance does.
  * overlap — ``degree`` = number of remote-KV stages (0 = one blocking stage, ``None`` = chosen by the planner:
    2 stages up to cp=2, 4 beyond, the values measured best on NVSwitch); stages are cut at equal mask area, which is what
    ``UniformOverlapAlg`` does in the reference; ``GreedyOverlapAlg`` and the cost factors are accepted and ignored.
"""
from __future__ import annotations

from dataclasses import dataclass, field

from .common.enum import AttnOverlapMode

__all__ = ["DistAttnConfig", "DispatchConfig", "DispatchAlg", "LBDispatchAlg", "DPDispatchAlg", "BSDispatchAlg",
           "MinHeapDispatchAlg", "ToppHeapDispatchAlg", "SequentialDispatchAlg", "SortedSequentialSelectAlg", "OverlapConfig",
           "OverlapAlg", "UniformOverlapAlg", "GreedyOverlapAlg", "GrpCollConfig"]


# ---- dispatch algorithms ---------------------------------------------------------------------------
@dataclass(frozen=True)
class DispatchAlg:
    """Marker base class; ``planner_name`` selects the implementation in meta/planner.py."""

    @property
    def planner_name(self) -> str:
        return "min_heap"


@dataclass(frozen=True)
class MinHeapDispatchAlg(DispatchAlg):
    pass


@dataclass(frozen=True)
class ToppHeapDispatchAlg(DispatchAlg):
    top_p: float = 0.0


@dataclass(frozen=True)
class LBDispatchAlg(DispatchAlg):
    pass


@dataclass(frozen=True)
class DPDispatchAlg(DispatchAlg):
    pass


@dataclass(frozen=True)
class BSDispatchAlg(DispatchAlg):
    pass


@dataclass(frozen=True)
class SequentialDispatchAlg(DispatchAlg):
    @property
    def planner_name(self) -> str:
        return "sequential"


@dataclass(frozen=True)
class SortedSequentialSelectAlg(DispatchAlg):
    allocation_ratio: float = 1.0

    @property
    def planner_name(self) -> str:
        return "sorted_sequential"


@dataclass(frozen=True)
class DispatchConfig:
    chunk_size: int | None = None  # None: taken from the key call / default 2048
    uneven_shard: bool = False
    alg: DispatchAlg | str = MinHeapDispatchAlg()

    def __post_init__(self):
        if self.uneven_shard:
            raise NotImplementedError("uneven_shard is not supported by the B200 planner (pad with compute_pad_size instead)")

    @property
    def planner_alg(self) -> str:
        return self.alg if isinstance(self.alg, str) else self.alg.planner_name


# ---- overlap ---------------------------------------------------------------------------------------
@dataclass(frozen=True)
class OverlapAlg:
    pass


@dataclass(frozen=True)
class UniformOverlapAlg(OverlapAlg):
    random_costs: bool = False
    random_seed: int | None = None


@dataclass(frozen=True)
class GreedyOverlapAlg(OverlapAlg):
    pass


@dataclass(frozen=True)
class OverlapConfig:
    enable: bool = True
    degree: int | None = 2  # number of remote-KV stages; 0 (or enable=False) => a single blocking stage; None => planner's choice
    mode: AttnOverlapMode = field(default=AttnOverlapMode.STATIC, kw_only=True)
    dynamic_max_degree: int | None = field(default=8, kw_only=True)
    min_chunk_size: int = field(default=512, kw_only=True)
    max_num_chunks: int = field(default=64, kw_only=True)
    alg: OverlapAlg = field(default=UniformOverlapAlg(), kw_only=True)
    calc_cost_factor: float = field(default=1.0, kw_only=True)
    comm_cost_factor: float = field(default=1.0, kw_only=True)

    @property
    def no_overlap(self) -> bool:
        return not self.enable or self.degree == 0

    def num_stages(self, cp_size: int) -> int:
        """Remote-KV stages the runtime will run for a cp group of this size."""
        if self.no_overlap or cp_size <= 1:
            return 1
        if self.degree is None or self.mode is AttnOverlapMode.DYNAMIC:
            d = 2 if cp_size <= 2 else 4
            return min(d, self.dynamic_max_degree) if self.dynamic_max_degree else d
        return max(1, self.degree)


@dataclass(frozen=True)
class GrpCollConfig:
    """Accepted for signature parity (reference: comm/primitive/grpcoll/_config.py); the B200 exchange is grouped NCCL P2P on
    row spans and has no SM / buffer knobs."""

    num_sms: int | None = None
    nvl_chunk_size: int | None = None
    nvl_buffer_size: int | None = None


@dataclass(frozen=True)
class DistAttnConfig:
    dispatch_config: DispatchConfig = field(default_factory=DispatchConfig)
    overlap_config: OverlapConfig = field(default_factory=OverlapConfig)
    grpcoll_config: GrpCollConfig = field(default_factory=GrpCollConfig)
    high_precision_reduce: bool = False  # fp32 wire dtype for the dK/dV group_reduce
