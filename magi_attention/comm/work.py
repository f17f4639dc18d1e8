from magiattention_b200.comm.work import GeneralWork, WorkWithPostProcessFn  # noqa: F401
