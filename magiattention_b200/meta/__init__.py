"""Host-side planning (mirror of the slice of magi_attention/meta the hot path consumes)."""
from .planner import (  # noqa: F401
    AttnArg,
    CalcMeta,
    CommMeta,
    DispatchMeta,
    GroupCollectiveArg,
    make_attn_meta,
    make_dispatch_meta,
    slice_area,
)
