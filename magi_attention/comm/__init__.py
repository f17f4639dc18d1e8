from magiattention_b200.comm import GeneralWork, WorkWithPostProcessFn, group_cast, group_reduce  # noqa: F401
