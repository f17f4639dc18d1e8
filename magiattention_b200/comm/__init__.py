"""Communication layer of the hot path (mirror of magi_attention/comm)."""
from .group_collective import group_cast, group_reduce  # noqa: F401
from .work import GeneralWork, WorkWithPostProcessFn  # noqa: F401
