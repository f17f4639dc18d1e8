"""DistAttnConfig — user-facing knobs (reference: magi_attention/config.py:54-71)."""
from __future__ import annotations

from dataclasses import dataclass, field


@dataclass(frozen=True)
class DispatchConfig:
    chunk_size: int | None = None  # None: taken from the key call / default 2048
    alg: str = "min_heap"


@dataclass(frozen=True)
class OverlapConfig:
    enable: bool = True
    degree: int | None = 2  # number of remote-KV stages; 0/None with enable=False => a single stage


@dataclass(frozen=True)
class DistAttnConfig:
    dispatch_config: DispatchConfig = field(default_factory=DispatchConfig)
    overlap_config: OverlapConfig = field(default_factory=OverlapConfig)
    high_precision_reduce: bool = False  # fp32 wire dtype for the dK/dV group_reduce
