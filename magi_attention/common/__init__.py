from magiattention_b200.common import AttnForwardMeta, AttnMaskType, AttnOverlapMode, AttnRange, AttnRanges, RangeError  # noqa: F401
