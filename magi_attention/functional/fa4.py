"""Alias of the reference's B200 kernel slot (magi_attention/functional/fa4.py): fa4_fwd / fa4_bwd / ffa_fa4_func."""
from magiattention_b200.functional.fa4 import fa4_bwd, fa4_fwd, is_fa4_installed  # noqa: F401
from magiattention_b200.functional.flex_flash_attn import ffa_fa4_func  # noqa: F401
