from . import grpcoll  # noqa: F401
