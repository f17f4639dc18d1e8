from magiattention_b200.comm import group_cast, group_reduce  # noqa: F401

__all__ = ["group_cast", "group_reduce"]
