"""Enums of the attention path.  Integer codes are the kernel's `attn_type_map` values
(reference: magi_attention/common/enum.py:58-76): 0 full, 1 causal, 2 inv-causal, 3 bi-causal."""
from __future__ import annotations

from enum import Enum


class AttnMaskType(Enum):
    FULL = "full"
    CAUSAL = "causal"
    BICAUSAL = "bi_causal"
    INVCAUSAL = "inv_causal"

    @classmethod
    def from_int_type(cls, int_type: int) -> "AttnMaskType":
        return _FROM_INT[int_type]

    def to_int_type(self) -> int:
        return _TO_INT[self]


_FROM_INT = {0: AttnMaskType.FULL, 1: AttnMaskType.CAUSAL, 2: AttnMaskType.INVCAUSAL, 3: AttnMaskType.BICAUSAL}
_TO_INT = {v: k for k, v in _FROM_INT.items()}


class AttnOverlapMode(Enum):
    STATIC = "static"
    DYNAMIC = "dynamic"


class AttnRole(Enum):
    """Which operand of the attention a tensor plays (reference: common/enum.py AttnRole)."""

    QUERY = "query"
    KEY = "key"
    VALUE = "value"
