from magiattention_b200.functional import (  # noqa: F401
    DistAttnFunc, DistAttnRuntime, FlexFlashAttnFunc, _flex_flash_attn_backward, _flex_flash_attn_forward,
    dispatch_func, dist_attn_func, flex_flash_attn_func, undispatch_func, roll_func, roll_simple_func, correct_attn_lse,
    correct_attn_lse_with_sink, correct_attn_out, correct_attn_out_lse, correct_attn_out_lse_with_sink,
    correct_attn_out_with_sink, ffa_fa4_func,
)
