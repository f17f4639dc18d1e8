from magiattention_b200.api import *  # noqa: F401,F403
from magiattention_b200.api import (  # noqa: F401
    calc_attn, clear_cache, compute_pad_size, dispatch, flex_flash_attn_func, get_most_recent_key, get_position_ids,
    infer_attn_mask_from_cu_seqlens, infer_attn_mask_from_sliding_window, infer_varlen_mask_from_batch,
    magi_attn_flex_dispatch, magi_attn_flex_key, magi_attn_varlen_dispatch, magi_attn_varlen_key, pad_at_dim,
    squash_batch_dim, undispatch, unpad_at_dim, make_flex_key_for_new_mask_after_dispatch,
    make_varlen_key_for_new_mask_after_dispatch, AttnForwardMeta, AttnMaskType, AttnOverlapMode, AttnRanges, DistAttnRuntimeKey,
    GeneralAttnMaskType, DistAttnConfig, DispatchConfig, OverlapConfig, roll, roll_simple, DistAttnRuntimeDictManager,
    dist_attn_runtime_dict_mgr,
)
