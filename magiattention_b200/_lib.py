"""ctypes binding of libb200_ffa.so (the C ABI in include/b200_ffa.h).

The library is the product; this module only marshals torch tensors into the plain-pointer
parameter blocks.  There is NO fallback: if the shared library is missing or a call fails, a
RuntimeError is raised (same error behaviour as the reference's TORCH_CHECK -> RuntimeError,
csrc/flexible_flash_attention/flex_flash_fwd.hpp:74-98).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200_FFA_LIB lets a developer point at an alternative build of the same library (kernel tuning variants)
LIB_PATH = os.environ.get("B200_FFA_LIB") or os.path.join(_HERE, "lib", "libb200_ffa.so")

B200_F16, B200_BF16, B200_F32 = 0, 1, 2
_DT = {torch.float16: B200_F16, torch.bfloat16: B200_BF16, torch.float32: B200_F32}


class FwdParams(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p),
        ("out", C.c_void_p), ("lse", C.c_void_p),
        ("q_ranges", C.c_void_p), ("k_ranges", C.c_void_p), ("attn_type_map", C.c_void_p),
        ("num_slices", C.c_int32),
        ("total_q", C.c_int32), ("total_k", C.c_int32), ("num_heads_q", C.c_int32), ("num_heads_kv", C.c_int32),
        ("head_dim", C.c_int32),
        ("q_stride_t", C.c_int64), ("q_stride_h", C.c_int64),
        ("k_stride_t", C.c_int64), ("k_stride_h", C.c_int64),
        ("v_stride_t", C.c_int64), ("v_stride_h", C.c_int64),
        ("o_stride_t", C.c_int64), ("o_stride_h", C.c_int64),
        ("lse_stride_t", C.c_int64), ("lse_stride_h", C.c_int64),
        ("softmax_scale", C.c_float), ("softcap", C.c_float),
        ("dtype", C.c_int32), ("out_dtype", C.c_int32),
        ("accumulate", C.c_int32), ("sm_margin", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("sink", C.c_void_p), ("seqlen_sink", C.c_int32), ("sink_layout", C.c_int32),
        ("max_logits", C.c_void_p),
    ]


class BwdParams(C.Structure):
    _fields_ = [
        ("dout", C.c_void_p), ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p),
        ("out", C.c_void_p), ("lse", C.c_void_p),
        ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p),
        ("q_ranges", C.c_void_p), ("k_ranges", C.c_void_p), ("attn_type_map", C.c_void_p),
        ("num_slices", C.c_int32),
        ("total_q", C.c_int32), ("total_k", C.c_int32), ("num_heads_q", C.c_int32), ("num_heads_kv", C.c_int32),
        ("head_dim", C.c_int32),
        ("q_stride_t", C.c_int64), ("q_stride_h", C.c_int64),
        ("k_stride_t", C.c_int64), ("k_stride_h", C.c_int64),
        ("v_stride_t", C.c_int64), ("v_stride_h", C.c_int64),
        ("o_stride_t", C.c_int64), ("o_stride_h", C.c_int64),
        ("do_stride_t", C.c_int64), ("do_stride_h", C.c_int64),
        ("dq_stride_t", C.c_int64), ("dq_stride_h", C.c_int64),
        ("dk_stride_t", C.c_int64), ("dk_stride_h", C.c_int64),
        ("dv_stride_t", C.c_int64), ("dv_stride_h", C.c_int64),
        ("lse_stride_t", C.c_int64), ("lse_stride_h", C.c_int64),
        ("softmax_scale", C.c_float), ("softcap", C.c_float),
        ("dtype", C.c_int32), ("dkv_dtype", C.c_int32),
        ("accumulate_dkv", C.c_int32), ("sm_margin", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("sink", C.c_void_p), ("dsink", C.c_void_p), ("seqlen_sink", C.c_int32), ("sink_layout", C.c_int32),
        ("row_stats", C.c_void_p), ("row_stats_ready", C.c_int32),
    ]


EXPORTS = [
    "b200_ffa_version", "b200_ffa_last_error", "b200_ffa_fwd_workspace_bytes", "b200_ffa_bwd_workspace_bytes",
    "b200_ffa_fwd", "b200_ffa_bwd", "b200_ffa_bwd_row_stats_bytes", "b200_lse_merge", "b200_sink_merge", "b200_range_reduce_sum", "b200_ffa_last_launch_count",
    "b200_nvl_alloc", "b200_nvl_free", "b200_nvl_export", "b200_nvl_import", "b200_nvl_release", "b200_nvl_copy_spans",
    "b200_nvl_signal", "b200_nvl_wait",
]

_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raise loudly if it is not built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} not found: build it with `python -m magiattention_b200.build` "
                        "(there is no CPU / eager fallback for the attention hot path)"
                    )
                l = C.CDLL(LIB_PATH)
                l.b200_ffa_version.restype = C.c_char_p
                l.b200_ffa_last_error.restype = C.c_char_p
                l.b200_ffa_fwd_workspace_bytes.restype = C.c_size_t
                l.b200_ffa_fwd_workspace_bytes.argtypes = [C.POINTER(FwdParams)]
                l.b200_ffa_bwd_workspace_bytes.restype = C.c_size_t
                l.b200_ffa_bwd_workspace_bytes.argtypes = [C.POINTER(BwdParams)]
                l.b200_ffa_bwd_row_stats_bytes.restype = C.c_size_t
                l.b200_ffa_bwd_row_stats_bytes.argtypes = [C.c_int32, C.c_int32]
                l.b200_ffa_fwd.restype = C.c_int
                l.b200_ffa_fwd.argtypes = [C.POINTER(FwdParams), C.c_void_p]
                l.b200_ffa_bwd.restype = C.c_int
                l.b200_ffa_bwd.argtypes = [C.POINTER(BwdParams), C.c_void_p]
                l.b200_lse_merge.restype = C.c_int
                l.b200_lse_merge.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                             C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
                l.b200_sink_merge.restype = C.c_int
                l.b200_sink_merge.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                              C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
                l.b200_range_reduce_sum.restype = C.c_int
                l.b200_range_reduce_sum.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                                    C.c_int32, C.c_int64, C.c_void_p]
                l.b200_ffa_last_launch_count.restype = C.c_int
                l.b200_nvl_alloc.restype = C.c_int
                l.b200_nvl_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
                l.b200_nvl_free.restype = C.c_int
                l.b200_nvl_free.argtypes = [C.c_void_p]
                l.b200_nvl_export.restype = C.c_int
                l.b200_nvl_export.argtypes = [C.c_void_p, C.c_char_p]
                l.b200_nvl_import.restype = C.c_int
                l.b200_nvl_import.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
                l.b200_nvl_release.restype = C.c_int
                l.b200_nvl_release.argtypes = [C.c_void_p]
                l.b200_nvl_copy_spans.restype = C.c_int
                l.b200_nvl_copy_spans.argtypes = [C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                                  C.POINTER(C.c_size_t), C.c_void_p]
                l.b200_nvl_signal.restype = C.c_int
                l.b200_nvl_signal.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_uint32, C.c_void_p]
                l.b200_nvl_wait.restype = C.c_int
                l.b200_nvl_wait.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
                _lib = l
    return _lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {lib().b200_ffa_last_error().decode()}")


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


_ws_cache: dict[tuple[int, int], torch.Tensor] = {}


def _workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    """Per-(device, stream) scratch buffer, grown on demand; stream-ordered reuse is safe."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream_ptr(device))
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def _req(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (the attention hot path has no CPU implementation)")
    if t.dim() != 3 or t.stride(2) != 1:
        raise RuntimeError(f"{name} must be [tokens, heads, head_dim] with contiguous head_dim")


def _ranges_ptr(r: torch.Tensor | None, n: int, name: str) -> int:
    if r is None:
        return 0
    if r.dtype != torch.int32 or not r.is_cuda or not r.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous int32 CUDA tensor")
    return r.data_ptr()


SINK_LAYOUTS = {"sh": 0, "ssh": 1}


def _sink_args(sink: torch.Tensor | None, sink_layout: str, total_q: int, heads: int, device) -> tuple[int, int, int]:
    """(pointer, seqlen_sink, layout code) of a sink tensor: fp32, "sh" [s_sink, hq] or "ssh" [tq, s_sink, hq]."""
    if sink is None:
        return 0, 0, 0
    if sink_layout == "shd":
        raise NotImplementedError(f"sink_layout {sink_layout} is not supported yet")
    if sink_layout not in SINK_LAYOUTS:
        raise ValueError(f"Invalid sink_layout {sink_layout}")
    if sink.dtype != torch.float32 or not sink.is_cuda or not sink.is_contiguous() or sink.device != device:
        raise RuntimeError("sink must be a contiguous fp32 CUDA tensor on q's device")
    want = (sink.shape[0], heads) if sink_layout == "sh" else (total_q, sink.shape[1] if sink.dim() == 3 else -1, heads)
    if tuple(sink.shape) != want:
        raise RuntimeError(f"sink with layout {sink_layout!r} must have shape {want}, got {tuple(sink.shape)}")
    return sink.data_ptr(), int(sink.shape[0] if sink_layout == "sh" else sink.shape[1]), SINK_LAYOUTS[sink_layout]


def ffa_fwd(
    q: torch.Tensor, k: torch.Tensor, v: torch.Tensor,
    out: torch.Tensor, lse: torch.Tensor,
    q_ranges: torch.Tensor, k_ranges: torch.Tensor, attn_type_map: torch.Tensor | None,
    softmax_scale: float, softcap: float = 0.0, accumulate: bool = False, sm_margin: int = 0,
    sink: torch.Tensor | None = None, sink_layout: str = "sh", max_logits: torch.Tensor | None = None,
) -> int:
    """Enqueue the forward on the current stream; returns the number of kernels launched."""
    l = lib()
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _req(t, n)
    if q.dtype not in (torch.float16, torch.bfloat16) or k.dtype != q.dtype or v.dtype != q.dtype:
        raise RuntimeError("q/k/v must share dtype float16 or bfloat16")
    if lse.dtype != torch.float32 or lse.dim() != 2:
        raise RuntimeError("lse must be fp32 [total_q, num_heads_q]")
    p = FwdParams()
    p.q, p.k, p.v, p.out, p.lse = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr()
    n = int(q_ranges.shape[0])
    p.q_ranges, p.k_ranges = _ranges_ptr(q_ranges, n, "q_ranges"), _ranges_ptr(k_ranges, n, "k_ranges")
    p.attn_type_map = _ranges_ptr(attn_type_map, n, "attn_type_map")
    p.num_slices = n
    p.total_q, p.total_k = q.shape[0], k.shape[0]
    p.num_heads_q, p.num_heads_kv, p.head_dim = q.shape[1], k.shape[1], q.shape[2]
    p.q_stride_t, p.q_stride_h = q.stride(0), q.stride(1)
    p.k_stride_t, p.k_stride_h = k.stride(0), k.stride(1)
    p.v_stride_t, p.v_stride_h = v.stride(0), v.stride(1)
    p.o_stride_t, p.o_stride_h = out.stride(0), out.stride(1)
    p.lse_stride_t, p.lse_stride_h = lse.stride(0), lse.stride(1)
    p.softmax_scale, p.softcap = softmax_scale, softcap
    p.dtype, p.out_dtype = _DT[q.dtype], _DT[out.dtype]
    p.accumulate, p.sm_margin = int(accumulate), sm_margin
    p.sink, p.seqlen_sink, p.sink_layout = _sink_args(sink, sink_layout, q.shape[0], q.shape[1], q.device)
    if max_logits is not None:
        if max_logits.dtype != torch.float32 or tuple(max_logits.shape) != (q.shape[1],) or not max_logits.is_contiguous() \
                or max_logits.device != q.device:
            raise RuntimeError("max_logits must be a contiguous fp32 [num_heads_q] tensor on q's device")
        p.max_logits = max_logits.data_ptr()
    nbytes = l.b200_ffa_fwd_workspace_bytes(C.byref(p))
    ws = _workspace(nbytes, q.device)
    p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
    _check(l.b200_ffa_fwd(C.byref(p), _stream_ptr(q.device)), "b200_ffa_fwd")
    return l.b200_ffa_last_launch_count()


def ffa_bwd(
    dout: torch.Tensor, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, lse: torch.Tensor,
    dq: torch.Tensor, dk: torch.Tensor, dv: torch.Tensor,
    q_ranges: torch.Tensor, k_ranges: torch.Tensor, attn_type_map: torch.Tensor | None,
    softmax_scale: float, softcap: float = 0.0, accumulate_dkv: bool = False, sm_margin: int = 0,
    sink: torch.Tensor | None = None, sink_layout: str = "sh", dsink: torch.Tensor | None = None,
    row_stats: "BwdRowStats | None" = None,
) -> int:
    """Enqueue the backward on the current stream (dq fp32 — and dsink when a sink is given — are accumulated into).

    `row_stats`: a BwdRowStats shared by the calls of ONE backward pass that use the same (out, dout, lse): the first call
    fills it, the later ones skip the preprocess kernel."""
    l = lib()
    for t, n in ((dout, "dout"), (q, "q"), (k, "k"), (v, "v"), (out, "out"), (dq, "dq"), (dk, "dk"), (dv, "dv")):
        _req(t, n)
    if dq.dtype != torch.float32:
        raise RuntimeError("dq must be fp32")
    if dk.dtype != dv.dtype or dk.dtype not in (q.dtype, torch.float32):
        raise RuntimeError("dk/dv must share a dtype equal to q's or fp32")
    if out.dtype != q.dtype or dout.dtype != q.dtype:
        raise RuntimeError("out/dout must have q's dtype")
    p = BwdParams()
    p.dout, p.q, p.k, p.v, p.out, p.lse = (t.data_ptr() for t in (dout, q, k, v, out, lse))
    p.dq, p.dk, p.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    n = int(q_ranges.shape[0])
    p.q_ranges, p.k_ranges = _ranges_ptr(q_ranges, n, "q_ranges"), _ranges_ptr(k_ranges, n, "k_ranges")
    p.attn_type_map = _ranges_ptr(attn_type_map, n, "attn_type_map")
    p.num_slices = n
    p.total_q, p.total_k = q.shape[0], k.shape[0]
    p.num_heads_q, p.num_heads_kv, p.head_dim = q.shape[1], k.shape[1], q.shape[2]
    for name, t in (("q", q), ("k", k), ("v", v), ("o", out), ("do", dout), ("dq", dq), ("dk", dk), ("dv", dv)):
        setattr(p, f"{name}_stride_t", t.stride(0))
        setattr(p, f"{name}_stride_h", t.stride(1))
    p.lse_stride_t, p.lse_stride_h = lse.stride(0), lse.stride(1)
    p.softmax_scale, p.softcap = softmax_scale, softcap
    p.dtype, p.dkv_dtype = _DT[q.dtype], _DT[dk.dtype]
    p.accumulate_dkv, p.sm_margin = int(accumulate_dkv), sm_margin
    p.sink, p.seqlen_sink, p.sink_layout = _sink_args(sink, sink_layout, q.shape[0], q.shape[1], q.device)
    if sink is not None:
        if dsink is None or dsink.dtype != torch.float32 or dsink.shape != sink.shape or not dsink.is_contiguous() \
                or dsink.device != q.device:
            raise RuntimeError("dsink must be a contiguous fp32 tensor shaped like sink")
        p.dsink = dsink.data_ptr()
    if row_stats is not None:
        p.row_stats, p.row_stats_ready = row_stats.bind(out, dout, lse), int(row_stats.ready)
    nbytes = l.b200_ffa_bwd_workspace_bytes(C.byref(p))
    ws = _workspace(nbytes, q.device)
    p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
    _check(l.b200_ffa_bwd(C.byref(p), _stream_ptr(q.device)), "b200_ffa_bwd")
    if row_stats is not None and q.shape[0] > 0 and k.shape[0] > 0:
        row_stats.ready = True
    return l.b200_ffa_last_launch_count()


class BwdRowStats:
    """The backward's per-(head, row) statistics (delta, lse * log2 e), kept between the stage calls of one pass.

    Bound to the first (out, dout, lse) it sees; a call with different tensors is a programming error and raises."""

    def __init__(self) -> None:
        self.buf: torch.Tensor | None = None
        self.ready = False
        self._key: tuple | None = None

    def bind(self, out: torch.Tensor, dout: torch.Tensor, lse: torch.Tensor) -> int:
        key = (out.data_ptr(), dout.data_ptr(), lse.data_ptr(), tuple(out.shape), torch.cuda.current_stream(out.device).cuda_stream)
        if self._key is None:
            self._key = key
            n = lib().b200_ffa_bwd_row_stats_bytes(out.shape[0], out.shape[1])
            self.buf = torch.empty(max(int(n), 256), dtype=torch.uint8, device=out.device)
        elif key != self._key:
            raise RuntimeError("BwdRowStats reused with different out / dout / lse tensors or on another stream")
        return self.buf.data_ptr()


def lse_merge(out: torch.Tensor, lse: torch.Tensor, out2: torch.Tensor, lse2: torch.Tensor) -> None:
    """(out, lse) <- LSE-merge((out, lse), (out2, lse2)) in place; all contiguous CUDA tensors."""
    l = lib()
    assert out.is_contiguous() and out2.is_contiguous() and lse.is_contiguous() and lse2.is_contiguous()
    rows, heads, dim = out.shape
    _check(l.b200_lse_merge(out.data_ptr(), lse.data_ptr(), out2.data_ptr(), lse2.data_ptr(), rows, heads, dim,
                            _DT[out.dtype], _DT[out2.dtype], _stream_ptr(out.device)), "b200_lse_merge")


def sink_merge(out: torch.Tensor, lse: torch.Tensor, sink: torch.Tensor, sink_layout: str = "sh") -> None:
    """(out, lse) <- merged with the attention sink, in place (out [rows, heads, dim] with contiguous dim, lse fp32)."""
    l = lib()
    _req(out, "out")
    if lse.dtype != torch.float32 or lse.shape != out.shape[:2] or not lse.is_cuda:
        raise RuntimeError("lse must be a CUDA fp32 [rows, heads] tensor")
    ptr, s_sink, layout = _sink_args(sink, sink_layout, out.shape[0], out.shape[1], out.device)
    _check(l.b200_sink_merge(out.data_ptr(), lse.data_ptr(), ptr, out.shape[0], out.shape[1], out.shape[2], s_sink, layout,
                             _DT[out.dtype], out.stride(0), out.stride(1), lse.stride(0), lse.stride(1),
                             _stream_ptr(out.device)), "b200_sink_merge")


def range_reduce_sum(dst: torch.Tensor, src: torch.Tensor, dst_ranges: torch.Tensor, src_ranges: torch.Tensor) -> None:
    """dst[dst_ranges[i]] += src[src_ranges[i]] for row spans (first dim); fp32 accumulate."""
    l = lib()
    assert dst.is_contiguous() and src.is_contiguous()
    row_elems = dst[0].numel() if dst.shape[0] else 0
    _check(l.b200_range_reduce_sum(dst.data_ptr(), _DT[dst.dtype], src.data_ptr(), _DT[src.dtype],
                                   dst_ranges.data_ptr(), src_ranges.data_ptr(), int(dst_ranges.shape[0]), row_elems,
                                   _stream_ptr(dst.device)), "b200_range_reduce_sum")


# ---- NVLink peer-memory exchange primitives (csrc/nvl_exchange.cu) -----------------------------------------
def nvl_alloc(nbytes: int) -> int:
    """Zero-filled cudaMalloc'd arena on the current device (IPC-exportable); returns the device pointer."""
    p = C.c_void_p()
    _check(lib().b200_nvl_alloc(C.byref(p), nbytes), "b200_nvl_alloc")
    return int(p.value)


def nvl_free(ptr: int) -> None:
    _check(lib().b200_nvl_free(ptr), "b200_nvl_free")


def nvl_export(ptr: int) -> bytes:
    buf = C.create_string_buffer(64)
    _check(lib().b200_nvl_export(ptr, buf), "b200_nvl_export")
    return buf.raw


def nvl_import(handle: bytes) -> int:
    p = C.c_void_p()
    _check(lib().b200_nvl_import(C.create_string_buffer(handle, 64), C.byref(p)), "b200_nvl_import")
    return int(p.value)


def nvl_release(ptr: int) -> None:
    _check(lib().b200_nvl_release(ptr), "b200_nvl_release")


class NvlCopyList:
    """A reusable (dst, src, bytes) span table marshalled once for b200_nvl_copy_spans."""

    def __init__(self, spans: list[tuple[int, int, int]]):
        n = len(spans)
        self.n = n
        self.dst = (C.c_void_p * max(n, 1))(*[s[0] for s in spans])
        self.src = (C.c_void_p * max(n, 1))(*[s[1] for s in spans])
        self.nbytes = (C.c_size_t * max(n, 1))(*[s[2] for s in spans])
        self.total = sum(s[2] for s in spans)

    def launch(self, stream: int) -> None:
        if self.n:
            _check(lib().b200_nvl_copy_spans(self.n, self.dst, self.src, self.nbytes, stream), "b200_nvl_copy_spans")


class NvlFlagList:
    """A reusable list of flag addresses for b200_nvl_signal / b200_nvl_wait (<= 64)."""

    def __init__(self, addrs: list[int]):
        assert len(addrs) <= 64
        self.n = len(addrs)
        self.arr = (C.c_void_p * max(self.n, 1))(*addrs)

    def signal(self, value: int, stream: int) -> None:
        if self.n:
            _check(lib().b200_nvl_signal(self.arr, self.n, value & 0xFFFFFFFF, stream), "b200_nvl_signal")

    def wait(self, value: int, err_ptr: int, timeout_ns: int, stream: int) -> None:
        if self.n:
            _check(lib().b200_nvl_wait(self.arr, self.n, value & 0xFFFFFFFF, err_ptr, timeout_ns, stream), "b200_nvl_wait")
