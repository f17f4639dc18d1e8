"""DistAttnConfig — user-facing knobs, same classes, field names, order and defaults as the reference
(magi_attention/config.py:54-71; DispatchConfig + *DispatchAlg: meta/solver/dispatch_solver.py:32-367;
OverlapConfig + *OverlapAlg: meta/solver/overlap_solver.py:26-150).

Everything here is frozen and hashable: the config is part of the runtime key.

What the B200 planner does with the knobs — a knob either works or raises, nothing is silently dropped:
  * dispatch algorithms — ``MinHeapDispatchAlg`` (default: chunks by decreasing mask area to the least-loaded rank, equal
    chunk counts), ``SequentialDispatchAlg`` (contiguous blocks of chunks per rank), ``SortedSequentialSelectAlg`` (area-sorted
    chunks dealt out in snake order) are implemented (meta/planner.py, pinned to the reference's partitions in
    tests/golden/dispatch_partitions.json).  ``LBDispatchAlg`` / ``DPDispatchAlg`` / ``BSDispatchAlg`` / ``ToppHeapDispatchAlg``
    raise NotImplementedError when used in a DispatchConfig.
  * overlap — ``degree`` = number of remote-KV stages (0 = no overlap: one blocking stage; ``None`` with
    ``mode=DYNAMIC`` = chosen by the planner: 2 stages up to cp=2, 4 beyond, capped by ``dynamic_max_degree``); stages are
    cut at equal mask area, which is what ``UniformOverlapAlg`` does in the reference (its ``random_costs`` only perturbs
    the reference's cost model and has no meaning for an area cut).  ``GreedyOverlapAlg`` raises.  ``min_chunk_size`` /
    ``max_num_chunks`` / ``calc_cost_factor`` / ``comm_cost_factor`` parameterise the reference's chunked cost model, which an
    exact area cut does not have: ``max_num_chunks`` is honoured as the upper bound of the stage count, the others are
    validated only.
  * ``uneven_shard`` raises (pad with compute_pad_size; SURVEY §8 marks it "next").
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
    """Base class; ``planner_name`` selects the implementation in meta/planner.py (None = not implemented)."""

    @property
    def planner_name(self) -> str | None:
        return None


@dataclass(frozen=True)
class MinHeapDispatchAlg(DispatchAlg):
    @property
    def planner_name(self) -> str:
        return "min_heap"


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


_PLANNER_ALGS = ("min_heap", "sequential", "sorted_sequential")


@dataclass(frozen=True)
class DispatchConfig:
    """Reference: meta/solver/dispatch_solver.py:359-367."""

    chunk_size: int | None = None
    uneven_shard: bool = False
    alg: DispatchAlg | str = MinHeapDispatchAlg()

    def __post_init__(self):
        if self.uneven_shard:
            raise NotImplementedError("uneven_shard is not supported by the B200 planner (pad with compute_pad_size instead)")
        name = self.alg if isinstance(self.alg, str) else self.alg.planner_name
        if name not in _PLANNER_ALGS:
            raise NotImplementedError(f"dispatch algorithm {self.alg!r} is not implemented by the B200 planner "
                                      f"(available: MinHeapDispatchAlg, SequentialDispatchAlg, SortedSequentialSelectAlg)")

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
    """Reference: meta/solver/overlap_solver.py:72-147 (same fields, order, defaults and validation).

    ``degree=0``: no overlap (one blocking stage); ``1``: local + 1 remote stage; ``N>=2``: local + N remote stages;
    ``None`` (with ``mode=DYNAMIC``): the planner picks."""

    mode: AttnOverlapMode = AttnOverlapMode.STATIC
    degree: int | None = 1
    dynamic_max_degree: int | None = 8
    min_chunk_size: int = 512
    max_num_chunks: int = 64
    alg: OverlapAlg = UniformOverlapAlg()
    calc_cost_factor: float = 1.0
    comm_cost_factor: float = 1.0
    _no_overlap: bool = field(init=False, repr=False, compare=False, default=False)

    @property
    def no_overlap(self) -> bool:
        return self._no_overlap

    @property
    def enable(self) -> bool:
        """Whether multi-stage overlap is enabled (degree >= 2 or dynamic)."""
        return not self._no_overlap and (self.degree is None or self.degree >= 2)

    def __post_init__(self):
        if isinstance(self.mode, bool):
            raise TypeError("OverlapConfig(mode, degree, ...): the first field is the AttnOverlapMode (reference layout); "
                            "use OverlapConfig(degree=N)")
        object.__setattr__(self, "_no_overlap", self.degree == 0)
        if self.degree is not None and self.degree <= 1:
            object.__setattr__(self, "mode", AttnOverlapMode.STATIC)
            if self._no_overlap:
                object.__setattr__(self, "degree", 1)
                object.__setattr__(self, "max_num_chunks", 1)
        if self.mode is AttnOverlapMode.STATIC:
            assert self.degree is not None, f"When using static overlap mode, the {self.degree=} should be set explicitly."
            assert self.degree <= self.max_num_chunks, f"The {self.max_num_chunks=} should be no less than {self.degree=}."
        elif self.mode is AttnOverlapMode.DYNAMIC:
            assert self.degree is None, f"When using dynamic overlap mode, the {self.degree=} should not be set."
            assert self.dynamic_max_degree is None or self.dynamic_max_degree > 0, (
                f"The {self.dynamic_max_degree=} should be greater than 0 if specified.")
        assert self.min_chunk_size > 0 and self.max_num_chunks > 0
        assert self.calc_cost_factor > 0.0 and self.comm_cost_factor > 0.0
        if not isinstance(self.alg, UniformOverlapAlg):
            raise NotImplementedError(f"overlap algorithm {self.alg!r} is not implemented by the B200 planner "
                                      "(stages are cut at equal mask area = UniformOverlapAlg)")

    def num_stages(self, cp_size: int) -> int:
        """Remote-KV stages the runtime will run for a cp group of this size."""
        if self._no_overlap or cp_size <= 1:
            return 1
        if self.degree is None:
            d = 2 if cp_size <= 2 else 4
            d = min(d, self.dynamic_max_degree) if self.dynamic_max_degree else d
        else:
            d = max(1, self.degree)
        return min(d, self.max_num_chunks)


@dataclass(frozen=True)
class GrpCollConfig:
    """Buffer / SM knobs of the reference's NVSHMEM native group collective (comm/primitive/grpcoll/_config.py).  The B200
    exchange moves rows with copy engines over NVLink peer memory and has none of them; the object is accepted so that
    configs written for the reference construct, and using the native collective itself raises
    (MAGI_ATTENTION_NATIVE_GRPCOLL=1, dist_attn_runtime_mgr.check_flag_comb)."""

    num_sms: int = 24
    nvl_chunk_size: int = 8
    nvl_buffer_size: int = 256
    rdma_chunk_size: int = 16
    rdma_buffer_size: int = 128
    num_nvl_bytes: int = int(2e9)
    num_rdma_bytes: int = 0


@dataclass(frozen=True)
class DistAttnConfig:
    """Reference: config.py:54-71 (dispatch_config, overlap_config, grpcoll_config); ``high_precision_reduce`` is this
    package's config-level spelling of MAGI_ATTENTION_BACKWARD_HIGH_PRECISION_REDUCE (fp32 wire dtype of the dK/dV reduce)."""

    dispatch_config: DispatchConfig = DispatchConfig()
    overlap_config: OverlapConfig = OverlapConfig()
    grpcoll_config: GrpCollConfig = GrpCollConfig()
    high_precision_reduce: bool = False
