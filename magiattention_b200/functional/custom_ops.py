"""The torch custom-op seam of the reference, on the B200 kernels.

``torch.ops.flex_flash_attn._flex_flash_attn_forward_compilable`` / ``_flex_flash_attn_backward_compilable`` keep the reference's
schema (magi_attention/functional/flex_flash_attn.py:193-330, 462-590): opaque, in-place (``mutates_args``) CUDA ops with no-op fake
implementations, so a ``torch.compile``d model can call the attention kernels without graph breaks.  Arguments that only steer the
reference's JIT kernel selection (block sizes, range-merge maps, sparse-load / index-attn tables, gqa packing, atomics switches) are
accepted and ignored; the ones the B200 backend does not implement (softcap, sparse_load, index_attn) raise.

Semantics follow the reference ops: ``out_`` / ``lse`` (and ``max_logits``) hold a previous partial result — zeros / -inf for a fresh
one — and are merged in place; ``dq`` / ``dk`` / ``dv`` (/ ``dsink``) are accumulated into (16-bit ``dk`` / ``dv`` are overwritten,
the kernels only accumulate into fp32).  The eager autograd path of this package (functional/flex_flash_attn.py) does not go through
these ops.
"""
from __future__ import annotations

import torch

from .. import _lib

__all__ = ["_flex_flash_attn_forward_compilable", "_flex_flash_attn_backward_compilable"]


def _check_supported(softcap: float, sparse_load: bool = False, index_attn: bool = False) -> None:
    if softcap != 0.0 or sparse_load or index_attn:
        raise RuntimeError("B200 FFA backend does not support softcap / sparse_load / index_attn")


@torch.library.custom_op("flex_flash_attn::_flex_flash_attn_forward_compilable", mutates_args=("out_", "lse", "max_logits"),
                         device_types="cuda")
def _flex_flash_attn_forward_compilable(
    q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, sink: torch.Tensor | None, sink_layout: str, out_: torch.Tensor,
    lse: torch.Tensor, q_ranges: torch.Tensor, k_ranges: torch.Tensor, attn_type_map: torch.Tensor | None, softmax_scale: float,
    softcap: float, out_type: torch.dtype | None, disable_fwd_atomic_reduction: bool, deterministic: bool, sm_margin: int,
    kblock_m: int | None, kblock_n: int | None, max_seqlen_q: int | None, auto_range_merge: bool,
    merge_q_ranges: torch.Tensor | None, fwd_qk_map: torch.Tensor | None, fwd_unique_count: torch.Tensor | None, swap_ab: bool,
    pack_gqa: bool, sparse_load: bool, sparse_load_loop_count: torch.Tensor | None, sparse_load_invalid_count: torch.Tensor | None,
    equal_k_range_size: torch.Tensor | None, index_attn: bool, index_attn_indices_2d: torch.Tensor | None, index_attn_max_topk: int,
    return_max_logits: bool, max_logits: torch.Tensor | None,
) -> None:
    _check_supported(softcap, sparse_load, index_attn)
    _lib.ffa_fwd(q, k, v, out_, lse, q_ranges, k_ranges, attn_type_map, softmax_scale, softcap, accumulate=True, sm_margin=sm_margin,
                 sink=sink, sink_layout=sink_layout, max_logits=max_logits if return_max_logits else None)


@_flex_flash_attn_forward_compilable.register_fake
def _(q, k, v, sink, sink_layout, out_, lse, q_ranges, k_ranges, attn_type_map, softmax_scale, softcap, out_type,
      disable_fwd_atomic_reduction, deterministic, sm_margin, kblock_m, kblock_n, max_seqlen_q, auto_range_merge, merge_q_ranges,
      fwd_qk_map, fwd_unique_count, swap_ab, pack_gqa, sparse_load, sparse_load_loop_count, sparse_load_invalid_count,
      equal_k_range_size, index_attn, index_attn_indices_2d, index_attn_max_topk, return_max_logits, max_logits) -> None:
    return None


@torch.library.custom_op("flex_flash_attn::_flex_flash_attn_backward_compilable", mutates_args=("dq", "dk", "dv", "dsink"),
                         device_types="cuda")
def _flex_flash_attn_backward_compilable(
    dout: torch.Tensor, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, sink: torch.Tensor | None, sink_layout: str,
    out_: torch.Tensor, lse: torch.Tensor, dq: torch.Tensor, dk: torch.Tensor, dv: torch.Tensor, dsink: torch.Tensor | None,
    q_ranges: torch.Tensor, k_ranges: torch.Tensor, attn_type_map: torch.Tensor | None, softmax_scale: float, softcap: float,
    dq_type: torch.dtype | None, dk_type: torch.dtype | None, dv_type: torch.dtype | None, disable_bwd_dkv_atomic_reduction: bool,
    deterministic: bool, sm_margin: int, auto_range_merge: bool, merge_k_ranges: torch.Tensor | None,
    bwd_kq_map: torch.Tensor | None, bwd_unique_count: torch.Tensor | None, swap_bwd_qk_loop: bool, pack_gqa: bool, cat_gqa: bool,
) -> None:
    _check_supported(softcap)
    out16 = out_ if out_.dtype == q.dtype else out_.to(q.dtype)
    _lib.ffa_bwd(dout, q, k, v, out16, lse, dq, dk, dv, q_ranges, k_ranges, attn_type_map, softmax_scale, softcap,
                 accumulate_dkv=dk.dtype == torch.float32, sm_margin=sm_margin, sink=sink, sink_layout=sink_layout, dsink=dsink)


@_flex_flash_attn_backward_compilable.register_fake
def _(dout, q, k, v, sink, sink_layout, out_, lse, dq, dk, dv, dsink, q_ranges, k_ranges, attn_type_map, softmax_scale, softcap,
      dq_type, dk_type, dv_type, disable_bwd_dkv_atomic_reduction, deterministic, sm_margin, auto_range_merge, merge_k_ranges,
      bwd_kq_map, bwd_unique_count, swap_bwd_qk_loop, pack_gqa, cat_gqa) -> None:
    return None
