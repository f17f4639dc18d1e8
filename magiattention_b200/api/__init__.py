"""User-facing API (mirror of magi_attention/api/__init__.py)."""
from ..functional.flex_flash_attn import flex_flash_attn_func  # noqa: F401
from .functools import (  # noqa: F401
    apply_padding,
    compute_pad_size,
    infer_attn_mask_from_cu_seqlens,
    infer_attn_mask_from_sliding_window,
    infer_varlen_mask_from_batch,
    pad_at_dim,
    squash_batch_dim,
    unpad_at_dim,
)
from .magi_attn_interface import (  # noqa: F401
    DistAttnRuntimeDictManager,
    dist_attn_runtime_dict_mgr,
    calc_attn,
    clear_cache,
    dispatch,
    get_most_recent_key,
    get_position_ids,
    get_runtime_mgr,
    magi_attn_flex_dispatch,
    magi_attn_flex_key,
    magi_attn_varlen_dispatch,
    magi_attn_varlen_key,
    make_flex_key_for_new_mask_after_dispatch,
    make_varlen_key_for_new_mask_after_dispatch,
    roll,
    roll_simple,
    undispatch,
)
# data structures, types and config classes used in the API signatures (reference api/__init__.py __all__)
from ..common.enum import AttnMaskType, AttnOverlapMode  # noqa: F401,E402
from ..common.forward_meta import AttnForwardMeta  # noqa: F401,E402
from ..common.ranges import AttnRanges  # noqa: F401,E402
from ..config import (  # noqa: F401,E402
    BSDispatchAlg, DispatchAlg, DispatchConfig, DistAttnConfig, DPDispatchAlg, GreedyOverlapAlg, GrpCollConfig, LBDispatchAlg,
    MinHeapDispatchAlg, OverlapAlg, OverlapConfig, SequentialDispatchAlg, SortedSequentialSelectAlg, ToppHeapDispatchAlg,
    UniformOverlapAlg,
)
from ..dist_attn_runtime_mgr import DistAttnRuntimeKey  # noqa: F401,E402

GeneralAttnMaskType = str | AttnMaskType | list  # str | AttnMaskType | list of either
