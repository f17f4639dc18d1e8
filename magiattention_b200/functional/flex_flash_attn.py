"""flex_flash_attn — the single-device attention op over AttnSlice metadata, B200-native.

Host-side mirror of the reference's operator interface for this path
(magi_attention/functional/flex_flash_attn.py): same names, argument meaning and error
behaviour for ``flex_flash_attn_func`` (:1066-1455), ``_flex_flash_attn_forward`` (:334-456) and
``_flex_flash_attn_backward`` (:594-693).  The compute goes through the C ABI of
``libb200_ffa.so`` (include/b200_ffa.h) — hand-written sm_100a kernels; there is no Triton,
no library GEMM and no CPU fallback.

Differences that are deliberate — a knob either works, is a pure scheduling hint of the reference's
Hopper kernel with no meaning here, or raises:
  * scheduling hints (results never depend on them in the reference either): ``disable_fwd_atomic_reduction`` /
    ``disable_bwd_dkv_atomic_reduction`` (promises that q / k ranges do not overlap — the kernels here never use global
    atomics for out / dk / dv), ``ref_block_size``, ``max_seqlen_q``, ``swap_ab``, ``pack_gqa``, ``cat_gqa``,
    ``swap_bwd_qk_loop``, ``auto_range_merge`` (tile shapes / loop orders / a range pre-merge of the sm90 kernel; this
    backend builds its own tile->slice plan) are accepted;
  * unsupported knobs raise like the reference's FA4 guard does (:1399-1421): ``softcap``, ``sparse_load``,
    ``index_attn`` / ``index_attn_indices``, sink layout ``"shd"``, ``deterministic=True`` (out / lse / dk / dv are
    bitwise reproducible, the fp32 dQ reduction order across kv tiles is not), ``sm_margin != 0`` (one-CTA-per-tile
    grids cannot leave SMs idle; the CP exchange here runs on copy engines and needs none).
Attention ``sink`` (fp32, layouts ``"sh"`` / ``"ssh"``, at most 8 sink tokens like the reference kernel) and
``return_max_logits`` are supported.
"""
from __future__ import annotations

import torch

from .. import _lib
from ..common.forward_meta import AttnForwardMeta

__all__ = ["flex_flash_attn_func", "ffa_fa4_func", "_flex_flash_attn_forward", "_flex_flash_attn_backward", "FlexFlashAttnFunc"]


DETERMINISTIC_SUPPORTED = False


def _reject_deterministic() -> None:
    raise NotImplementedError("deterministic=True: out, lse, dk and dv of the B200 kernels are bitwise reproducible, but dq is "
                              "reduced across kv tiles with TMA reduce-adds whose order is not fixed")


def _reject_sm_margin(sm_margin: int) -> None:
    if sm_margin:
        raise NotImplementedError("sm_margin != 0 is not supported: the B200 kernels are one-CTA-per-tile grids that cannot "
                                  "leave SMs idle (the CP exchange runs on copy engines and needs no SM reservation)")


def _maybe_contiguous(x: torch.Tensor | None) -> torch.Tensor | None:
    return x.contiguous() if x is not None and not x.is_contiguous() else x


def _check_ranges(q_ranges: torch.Tensor, k_ranges: torch.Tensor, attn_type_map: torch.Tensor | None) -> None:
    if q_ranges.dim() != 2 or q_ranges.shape[1] != 2 or q_ranges.shape != k_ranges.shape:
        raise RuntimeError("q_ranges / k_ranges must both be [num_slices, 2]")
    if q_ranges.dtype != torch.int32 or k_ranges.dtype != torch.int32:
        raise RuntimeError("q_ranges / k_ranges must be int32")
    if attn_type_map is not None and (attn_type_map.dtype != torch.int32 or attn_type_map.shape[0] != q_ranges.shape[0]):
        raise RuntimeError("attn_type_map must be int32 [num_slices]")


def _flex_flash_attn_forward(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    sink: torch.Tensor | None,
    sink_layout: str,
    out: torch.Tensor | None,
    lse: torch.Tensor | None,
    q_ranges: torch.Tensor,
    k_ranges: torch.Tensor,
    attn_type_map: torch.Tensor | None,
    softmax_scale: float,
    softcap: float,
    out_type: torch.dtype | None,
    disable_fwd_atomic_reduction: bool = False,
    deterministic: bool = False,
    sm_margin: int = 0,
    **unused,
) -> tuple[torch.Tensor, AttnForwardMeta]:
    """Raw forward launcher with ACCUMULATOR semantics: when ``out``/``lse`` are given they hold a
    previous partial result (over a disjoint key set) and are LSE-merged in place with this call's
    result inside the kernel epilogue (reference: dist_attn.py:1296-1297; pinned by
    tests/test_attn/test_flex_flash_attn.py:362-540)."""
    if unused.get("sparse_load") or unused.get("index_attn"):
        raise RuntimeError("B200 FFA backend does not support sparse_load / index_attn")
    if deterministic:
        _reject_deterministic()
    _reject_sm_margin(sm_margin)
    return_max_logits = bool(unused.get("return_max_logits", False))
    max_logits = unused.get("max_logits")
    if return_max_logits and max_logits is None:
        # reference: flex_flash_attn.py:397-403 — a fresh buffer starts at -inf, a given one is max-reduced into
        max_logits = torch.full((q.size(1),), float("-inf"), dtype=torch.float32, device=q.device)
    q, k, v, sink, q_ranges, k_ranges = (_maybe_contiguous(x) for x in (q, k, v, sink, q_ranges, k_ranges))
    _check_ranges(q_ranges, k_ranges, attn_type_map)
    accumulate = out is not None
    if (out is None) != (lse is None):
        raise RuntimeError("out and lse must be given together")
    if out is None:
        out = torch.empty_like(q, dtype=out_type or q.dtype)
        lse = torch.empty((q.size(0), q.size(1)), dtype=torch.float32, device=q.device)
    _lib.ffa_fwd(q, k, v, out, lse, q_ranges, k_ranges, attn_type_map, softmax_scale, softcap,
                 accumulate=accumulate, sm_margin=sm_margin, sink=sink, sink_layout=sink_layout,
                 max_logits=max_logits if return_max_logits else None)
    return out, AttnForwardMeta(lse=lse, max_logits=max_logits if return_max_logits else None)


def _flex_flash_attn_backward(
    dout: torch.Tensor,
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    sink: torch.Tensor | None,
    sink_layout: str,
    out: torch.Tensor,
    lse: torch.Tensor,
    dq: torch.Tensor | None,
    dk: torch.Tensor | None,
    dv: torch.Tensor | None,
    dsink: torch.Tensor | None,
    q_ranges: torch.Tensor,
    k_ranges: torch.Tensor,
    attn_type_map: torch.Tensor | None,
    softmax_scale: float,
    softcap: float,
    dq_type: torch.dtype | None = None,
    dk_type: torch.dtype | None = None,
    dv_type: torch.dtype | None = None,
    disable_bwd_dkv_atomic_reduction: bool = False,
    deterministic: bool = False,
    sm_margin: int = 0,
    row_stats: "_lib.BwdRowStats | None" = None,
    **unused,
) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor | None]:
    """Raw backward launcher.  ``out``/``lse`` are the GLOBAL forward results, so partial gradients over
    disjoint key sets add exactly.  A given fp32 ``dq`` is accumulated into; given fp32 ``dk``/``dv`` are
    accumulated into as well (reference: flex_flash_attn.py:637-647, dist_attn.py:1432).  With a ``sink`` the
    fp32 ``dsink`` (zero-initialised when not given, :652-657) is accumulated into and returned."""
    if deterministic:
        _reject_deterministic()
    _reject_sm_margin(sm_margin)
    dout, q, k, v, sink, out, q_ranges, k_ranges = (
        _maybe_contiguous(x) for x in (dout, q, k, v, sink, out, q_ranges, k_ranges))
    if sink is not None and dsink is None:
        dsink = torch.zeros_like(sink, dtype=torch.float32)
    _check_ranges(q_ranges, k_ranges, attn_type_map)
    if out.dtype != q.dtype:
        out = out.to(q.dtype)
    if dq is None:
        dq = torch.zeros_like(q, dtype=torch.float32)
    elif dq.dtype != torch.float32:
        raise RuntimeError("dq accumulator must be fp32")
    if (dk is None) != (dv is None):
        raise RuntimeError("dk and dv must be given together")
    accumulate_dkv = dk is not None
    if dk is None:
        dk = torch.empty_like(k, dtype=dk_type or torch.float32)
        dv = torch.empty_like(v, dtype=dv_type or torch.float32)
    _lib.ffa_bwd(dout, q, k, v, out, lse, dq, dk, dv, q_ranges, k_ranges, attn_type_map, softmax_scale, softcap,
                 accumulate_dkv=accumulate_dkv, sm_margin=sm_margin, sink=sink, sink_layout=sink_layout,
                 dsink=dsink if sink is not None else None, row_stats=row_stats)
    return dq, dk, dv, (dsink if sink is not None else None)


class FlexFlashAttnFunc(torch.autograd.Function):
    """Autograd glue (reference: FlexFlashAttnFunc, flex_flash_attn.py:699-1060)."""

    @staticmethod
    def forward(ctx, q, k, v, sink, sink_layout, q_ranges, k_ranges, attn_type_map, softmax_scale, softcap, sm_margin,
                return_max_logits):
        out, meta = _flex_flash_attn_forward(
            q, k, v, sink, sink_layout, None, None, q_ranges, k_ranges, attn_type_map, softmax_scale, softcap,
            out_type=q.dtype, sm_margin=sm_margin, return_max_logits=return_max_logits,
        )
        ctx.save_for_backward(q, k, v, sink, out, meta.lse, q_ranges, k_ranges, attn_type_map)
        ctx.softmax_scale, ctx.softcap, ctx.sm_margin, ctx.sink_layout = softmax_scale, softcap, sm_margin, sink_layout
        ctx.mark_non_differentiable(meta.lse)
        if meta.max_logits is not None:
            ctx.mark_non_differentiable(meta.max_logits)
        return out, meta.lse, meta.max_logits

    @staticmethod
    def backward(ctx, dout, *args):
        q, k, v, sink, out, lse, q_ranges, k_ranges, attn_type_map = ctx.saved_tensors
        dq, dk, dv, dsink = _flex_flash_attn_backward(
            dout, q, k, v, sink, ctx.sink_layout, out, lse, None, None, None, None, q_ranges, k_ranges, attn_type_map,
            ctx.softmax_scale, ctx.softcap, dk_type=k.dtype, dv_type=v.dtype, sm_margin=ctx.sm_margin,
        )
        if dsink is not None:
            dsink = dsink.to(sink.dtype)
        return dq.to(q.dtype), dk, dv, dsink, None, None, None, None, None, None, None, None


def flex_flash_attn_func(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    q_ranges: torch.Tensor | None = None,
    k_ranges: torch.Tensor | None = None,
    attn_type_map: torch.Tensor | None = None,
    *,
    index_attn_indices: torch.Tensor | None = None,
    q_block_size: int = 1,
    k_block_size: int = 1,
    sink: torch.Tensor | None = None,
    sink_layout: str = "sh",
    softmax_scale: float | None = None,
    softcap: float = 0.0,
    deterministic: bool = False,
    sm_margin: int = 0,
    disable_fwd_atomic_reduction: bool = False,
    disable_bwd_dkv_atomic_reduction: bool = False,
    ref_block_size: tuple[int, int] | None = None,
    max_seqlen_q: int | None = None,
    auto_range_merge: bool = False,
    swap_ab: bool = False,
    pack_gqa: bool = False,
    cat_gqa: bool = False,
    sparse_load: bool = False,
    index_attn: bool = False,
    swap_bwd_qk_loop: bool = False,
    return_max_logits: bool = False,
) -> tuple[torch.Tensor, AttnForwardMeta]:
    """Attention of ``q`` [tq, hq, d] over ``k``/``v`` [tk, hk, d] restricted to the union of AttnSlices
    ``(q_ranges[i], k_ranges[i], attn_type_map[i])`` (int32 device tensors; types 0 full, 1 causal,
    2 inv-causal, 3 bi-causal).  Returns ``(out [tq,hq,d] in q's dtype, AttnForwardMeta(lse [tq,hq] fp32))``;
    rows covered by no slice get out = 0, lse = -inf.  Same signature as the reference
    (functional/flex_flash_attn.py:1066-1095)."""
    assert (q_ranges is not None) != (index_attn_indices is not None), (
        "Exactly one of (q_ranges + k_ranges) or index_attn_indices must be provided."
    )
    bad = [name for name, on in {
        "index_attn_indices": index_attn_indices is not None, "softcap": softcap != 0.0,
        "sparse_load": sparse_load, "index_attn": index_attn, "deterministic": deterministic, "sm_margin": sm_margin != 0,
    }.items() if on]
    assert not bad, f"B200 FFA backend does not support the following features: {bad}."
    assert k_ranges is not None, "k_ranges must be provided together with q_ranges"
    assert not (swap_bwd_qk_loop and deterministic), "Deterministic mode is not supported when swap_bwd_qk_loop is enabled."
    assert q.dim() == 3 and k.dim() == 3 and v.dim() == 3, "q/k/v must be [tokens, heads, head_dim]"
    if softmax_scale is None:
        softmax_scale = q.shape[-1] ** (-0.5)
    if return_max_logits:
        assert q.size(1) <= 128, f"num_qheads ({q.size(1)}) must be <= 128 when return_max_logits is set"  # reference :404-409
    out, lse, max_logits = FlexFlashAttnFunc.apply(
        q, k, v, sink, sink_layout, q_ranges, k_ranges, attn_type_map, softmax_scale, softcap, sm_margin, return_max_logits)
    return out, AttnForwardMeta(lse=lse, max_logits=max_logits)


def ffa_fa4_func(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, q_ranges: torch.Tensor, k_ranges: torch.Tensor,
                 attn_type_map: torch.Tensor | None = None, *, softmax_scale: float | None = None, softcap: float = 0.0,
                 deterministic: bool = False, reuse_attn_arg: bool = False) -> tuple[torch.Tensor, torch.Tensor]:
    """The reference's B200 incumbent entry point (functional/fa4.py:421-470, used by exps/attn/run_blackwell_benchmark.py):
    same kernels as :func:`flex_flash_attn_func`, returns ``(out, lse)``.  ``reuse_attn_arg`` has nothing to reuse here —
    slices are consumed directly, there is no per-call mask conversion."""
    out, meta = flex_flash_attn_func(q, k, v, q_ranges, k_ranges, attn_type_map, softmax_scale=softmax_scale, softcap=softcap,
                                     deterministic=deterministic)
    return out, meta.lse
