"""Import stub: the reference imports `debugpy` at magi_attention/utils/debug.py:17 (a dev-only remote debugger that is
not in this image and cannot be installed offline).  Nothing on the timed path calls it."""


def listen(*a, **k):
    raise RuntimeError("debugpy is not available in this image")


def wait_for_client(*a, **k):
    raise RuntimeError("debugpy is not available in this image")
