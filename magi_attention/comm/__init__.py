from magiattention_b200.comm import GeneralWork, WorkWithPostProcessFn, group_cast, group_reduce  # noqa: F401

from . import primitive  # noqa: F401,E402
