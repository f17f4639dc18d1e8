"""magiattention_b200 — B200-native context-parallel attention engine behind the MagiAttention API.

Hot path only: flex-flash-attention fwd/bwd as hand-written sm_100a kernels behind a
C ABI (include/b200_ffa.h), the group-cast / group-reduce KV exchange over NCCL P2P, and the autograd
glue + runtime entry points that call them.
"""
__version__ = "0.2.0"

from . import common  # noqa: F401,E402
