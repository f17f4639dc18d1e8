"""Autograd glue of the attention hot path (mirror of magi_attention/functional)."""
from .flex_flash_attn import (  # noqa: F401
    FlexFlashAttnFunc,
    _flex_flash_attn_backward,
    _flex_flash_attn_forward,
    flex_flash_attn_func,
)
from .dist_attn import DistAttnFunc, DistAttnRuntime, dist_attn_func  # noqa: F401,E402
from .dispatch import dispatch_func, undispatch_func  # noqa: F401,E402
