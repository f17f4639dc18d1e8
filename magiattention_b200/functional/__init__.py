"""Autograd glue of the attention hot path (mirror of magi_attention/functional)."""
from .flex_flash_attn import (  # noqa: F401
    FlexFlashAttnFunc,
    _flex_flash_attn_backward,
    _flex_flash_attn_forward,
    ffa_fa4_func,
    flex_flash_attn_func,
)
from .dist_attn import DistAttnFunc, DistAttnRuntime, dist_attn_func  # noqa: F401,E402
from .dispatch import dispatch_func, undispatch_func  # noqa: F401,E402
from .roll import roll_func, roll_simple_func  # noqa: F401,E402
from .utils import (  # noqa: F401,E402
    calc_lse_sink,
    correct_attn_lse,
    correct_attn_lse_with_sink,
    correct_attn_out,
    correct_attn_out_lse,
    correct_attn_out_lse_with_sink,
    correct_attn_out_with_sink,
)

try:  # registers torch.ops.flex_flash_attn.* (the reference's torch.compile seam); skipped if the namespace already has them
    from . import custom_ops  # noqa: F401,E402
except RuntimeError:  # pragma: no cover - the ops were registered by another library in this process
    custom_ops = None

from .fa4 import fa4_bwd, fa4_fwd  # noqa: F401,E402
