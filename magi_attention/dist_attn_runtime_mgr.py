from magiattention_b200.dist_attn_runtime_mgr import (  # noqa: F401
    DistAttnRuntimeKey, DistAttnRuntimeMgr, init_dist_attn_runtime_key, init_dist_attn_runtime_mgr,
)
