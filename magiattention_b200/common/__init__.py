"""Host-side value types of the attention path (mirror of magi_attention/common for the hot path):
AttnRange / AttnRanges (token spans), AttnMaskType (slice mask kinds), AttnForwardMeta (lse, max_logits)."""
from .enum import AttnMaskType, AttnOverlapMode, AttnRole  # noqa: F401
from .forward_meta import AttnForwardMeta  # noqa: F401
from .range import AttnRange, RangeError  # noqa: F401
from .ranges import AttnRanges  # noqa: F401
