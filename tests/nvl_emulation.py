"""CPU emulation of the NVLink peer-memory primitives — TEST INFRASTRUCTURE ONLY (never imported by the product).

`comm/nvl_exchange.py` is pure host logic on top of eight C-ABI primitives (b200_nvl_alloc / free / export / import /
release / copy_spans / signal / wait, include/b200_ffa.h) plus CUDA streams and events.  Installed inside a gloo worker
process, this module replaces exactly those primitives with host equivalents so the REAL exchange code (arena layout,
pull tables, slot rotation, flag protocol, host order of signals and waits) runs on CPU across real processes:

  arena            a MAP_SHARED file mapping (one per rank); "export" = its path, "import" = mapping the peer's file
  copy_spans       memmove between real addresses (tensor storage or mapped arenas)
  signal / wait    stores / polls of the 32-bit epoch words inside the mapped arenas, same wrap-safe comparison and the
                   same timeout -> error word behaviour as csrc/nvl_exchange.cu (the test turns the error into a raise)
  streams, events  every enqueued operation executes immediately in host order, a wait blocks the host

That is the most restrictive legal device schedule (one in-order queue per rank following host order): if a pass does
not deadlock here it cannot deadlock with concurrent streams, which only remove ordering; and it is the schedule that
implicit host/device synchronisation (a cudaMalloc inside the caching allocator) forces on the real device, the case
the host-order rule in NvlKVExchange.begin_reduce exists for.  Cross-rank interleavings are real (separate processes,
optional per-rank skew)."""
import contextlib
import ctypes
import mmap
import os
import random
import time

import torch


class _Event:
    def wait(self, *a):
        pass

    def synchronize(self):
        pass

    def query(self):
        return True


class _Stream:
    _n = 0

    def __init__(self, device=None, priority=0):
        _Stream._n += 1
        self.cuda_stream = _Stream._n

    def wait_stream(self, other):
        pass

    def wait_event(self, ev):
        pass

    def record_event(self, ev=None):
        return _Event()

    def synchronize(self):
        pass


class _Arena:
    def __init__(self, path: str, nbytes: int, create: bool):
        self.path, self.owner = path, create
        fd = os.open(path, os.O_RDWR | (os.O_CREAT | os.O_EXCL if create else 0), 0o600)
        try:
            if create:
                os.ftruncate(fd, nbytes)  # zero-filled, like b200_nvl_alloc
            self.mm = mmap.mmap(fd, nbytes if create else os.fstat(fd).st_size)
        finally:
            os.close(fd)
        self.anchor = ctypes.c_char.from_buffer(self.mm)
        self.base = ctypes.addressof(self.anchor)

    def close(self):
        self.anchor = None
        try:
            self.mm.close()
        except BufferError:  # a tensor view is still alive: the mapping goes with the process
            pass
        if self.owner:
            with contextlib.suppress(FileNotFoundError):
                os.unlink(self.path)


_state = {"dir": None, "n": 0, "arenas": {}, "skew_s": 0.0, "log": []}


def _alloc(nbytes: int) -> int:
    _state["n"] += 1
    a = _Arena(os.path.join(_state["dir"], f"arena_{os.getpid()}_{_state['n']}"), nbytes, True)
    _state["arenas"][a.base] = a
    return a.base


def _export(base: int) -> bytes:
    return _state["arenas"][base].path.encode()


def _import(handle: bytes) -> int:
    a = _Arena(handle.decode(), 0, False)
    _state["arenas"][a.base] = a
    return a.base


def _release(base: int) -> None:
    _state["arenas"].pop(base).close()


class _CopyList:
    def __init__(self, spans):
        self.spans = list(spans)
        self.n = len(self.spans)
        self.total = sum(s[2] for s in self.spans)

    def launch(self, stream: int) -> None:
        if _state["skew_s"]:
            time.sleep(random.random() * _state["skew_s"])
        for dst, src, n in self.spans:
            ctypes.memmove(dst, src, n)
        mine = [a for a in _state["arenas"].values() if a.owner]
        publish = any(a.base <= s[0] < a.base + len(a.mm) for a in mine for s in self.spans)  # writes this rank's own arena
        _state["log"].append(("publish" if publish else "copy", self.n, self.total))


class _FlagList:
    def __init__(self, addrs):
        assert len(addrs) <= 64
        self.words = [ctypes.c_uint32.from_address(a) for a in addrs]
        self.n = len(self.words)

    def signal(self, value: int, stream: int) -> None:
        for w in self.words:
            w.value = value & 0xFFFFFFFF
        _state["log"].append(("signal", self.n, value))

    def wait(self, value: int, err_ptr: int, timeout_ns: int, stream: int) -> None:
        t_end = time.monotonic() + timeout_ns * 1e-9
        v = value & 0xFFFFFFFF
        while True:
            if all(((w.value - v) & 0xFFFFFFFF) < 0x80000000 for w in self.words):  # flag >= value, wrap-safe
                _state["log"].append(("wait", self.n, value))
                return
            if time.monotonic() > t_end:
                ctypes.c_int32.from_address(err_ptr).value = 1
                raise TimeoutError(f"emulated b200_nvl_wait: flags {[w.value for w in self.words]} never reached {value}")
            time.sleep(0.0002)


def _range_reduce_sum(dst, src, dst_ranges, src_ranges) -> None:
    for (d0, d1), (s0, s1) in zip(dst_ranges.tolist(), src_ranges.tolist()):
        dst[d0:d1] += src[s0:s1].to(dst.dtype)


def install(shared_dir: str, group, skew_s: float = 0.0, timeout_s: float = 20.0):
    """Patch this PROCESS: NVLink primitives -> host emulation, CUDA streams -> immediate execution, and register the
    emulated arena domain for ``group``.  Returns the domain."""
    import magiattention_b200._lib as _lib
    from magiattention_b200.comm import nvl_exchange as nx

    _state.update(dir=shared_dir, skew_s=skew_s)
    random.seed(os.getpid())
    _lib.nvl_alloc, _lib.nvl_free, _lib.nvl_export, _lib.nvl_import, _lib.nvl_release = _alloc, _release, _export, _import, _release
    _lib.NvlCopyList, _lib.NvlFlagList, _lib.range_reduce_sum = _CopyList, _FlagList, _range_reduce_sum
    cur = _Stream()
    torch.cuda.Stream = _Stream
    torch.cuda.current_stream = lambda device=None: cur
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.synchronize = lambda device=None: None
    torch.cuda.current_device = lambda: 0
    nx.nvl_available = lambda g: True
    nx.NvlDomain.TIMEOUT_NS = int(timeout_s * 1e9)

    class EmuDomain(nx.NvlDomain):
        def __init__(self, g):
            pin = torch.Tensor.pin_memory
            torch.Tensor.pin_memory = lambda t, *a, **k: t
            try:
                super().__init__(g)
            finally:
                torch.Tensor.pin_memory = pin
            self.device = torch.device("cpu")

        def _map_view(self, total):
            return torch.frombuffer(_state["arenas"][self.base].mm, dtype=torch.uint8, count=total)

    dom = EmuDomain(group)
    nx.NvlDomain._domains[id(group)] = dom
    return dom


def op_log() -> list:
    return _state["log"]
