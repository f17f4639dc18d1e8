from magiattention_b200.common import AttnForwardMeta, AttnMaskType, AttnOverlapMode, AttnRange, AttnRanges, AttnRole, RangeError  # noqa: F401
