from magiattention_b200.config import DispatchConfig, DistAttnConfig, OverlapConfig  # noqa: F401
