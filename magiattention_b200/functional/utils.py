"""LSE-correction helpers of online / staged attention — the reference's magi_attention/functional/utils.py surface.

``correct_attn_out_lse`` (:467-558, a Triton kernel there) and ``correct_attn_out_lse_with_sink`` (:634-672) run the C ABI's
``b200_lse_merge`` / ``b200_sink_merge`` kernels; the single-tensor variants are the same few elementwise torch ops the
reference uses (:286-358, :561-631).  The CP runtime itself needs none of them — stages merge in the forward kernel's
epilogue — they exist for callers that hold partial results of their own.
"""
from __future__ import annotations

import torch

from .. import _lib

__all__ = ["correct_attn_lse", "correct_attn_out", "correct_attn_out_lse", "calc_lse_sink", "correct_attn_lse_with_sink",
           "correct_attn_out_with_sink", "correct_attn_out_lse_with_sink"]

_NEG_INF = float("-inf")


def _safe_sub(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a - b with (-inf) - (-inf) := -inf (reference safe_subtract :25-35)."""
    return (a - b).masked_fill((a == _NEG_INF) & (b == _NEG_INF), _NEG_INF)


def correct_attn_lse(lse1: torch.Tensor, lse2: torch.Tensor, inplace: bool = False) -> torch.Tensor:
    """lse = max + log1p(exp(min - max)), (-inf, -inf) -> -inf (reference :286-319)."""
    mx, mn = torch.maximum(lse1, lse2), torch.minimum(lse1, lse2)
    lse = mx + torch.log1p(torch.exp(_safe_sub(mn, mx)))
    lse = lse.masked_fill(mx == _NEG_INF, _NEG_INF)
    return lse1.copy_(lse) if inplace else lse


def correct_attn_out(out1: torch.Tensor, lse1: torch.Tensor, out2: torch.Tensor, lse2: torch.Tensor, lse: torch.Tensor,
                     inplace: bool = False) -> torch.Tensor:
    """out = exp(lse1 - lse) out1 + exp(lse2 - lse) out2 given the merged ``lse`` (reference :322-358)."""
    assert out1.dtype == out2.dtype and lse1.dtype == lse2.dtype == lse.dtype
    w1, w2 = _safe_sub(lse1, lse).exp().unsqueeze(-1), _safe_sub(lse2, lse).exp().unsqueeze(-1)
    out = w1 * out1 + w2 * out2
    return out1.copy_(out) if inplace else out.to(out1.dtype)


def correct_attn_out_lse(out1: torch.Tensor, lse1: torch.Tensor, out2: torch.Tensor, lse2: torch.Tensor,
                         inplace: bool = False) -> tuple[torch.Tensor, torch.Tensor]:
    """LSE-merge of two partial attention results over disjoint key sets, one fused kernel (``b200_lse_merge``)."""
    if not inplace:
        out1, lse1 = out1.clone(), lse1.clone()
    if not (out1.is_contiguous() and lse1.is_contiguous()):
        raise RuntimeError("correct_attn_out_lse(inplace=True) needs contiguous out1 / lse1")
    _lib.lse_merge(out1, lse1, out2.contiguous(), lse2.contiguous())
    return out1, lse1


def calc_lse_sink(sink: torch.Tensor, seqlen_q: int, sink_layout: str = "sh") -> torch.Tensor:
    """log-sum-exp of the sink logits per (row, head): [seqlen_q, num_heads_q] (reference :234-280)."""
    if sink_layout == "sh":
        assert sink.dim() == 2
        return torch.logsumexp(sink, dim=0, keepdim=True).repeat(seqlen_q, 1)
    if sink_layout == "ssh":
        assert sink.dim() == 3 and sink.size(0) == seqlen_q
        return torch.logsumexp(sink, dim=1)
    if sink_layout == "shd":
        raise NotImplementedError(f"{sink_layout} is not supported yet")
    raise ValueError(f"Invalid sink_layout: {sink_layout}")


def correct_attn_lse_with_sink(lse: torch.Tensor, sink: torch.Tensor, sink_layout: str = "sh", inplace: bool = False) -> torch.Tensor:
    """lse merged with the sink's lse (reference :561-590)."""
    return correct_attn_lse(lse, calc_lse_sink(sink, lse.size(0), sink_layout).to(lse.dtype), inplace=inplace)


def correct_attn_out_with_sink(out: torch.Tensor, lse: torch.Tensor, sink: torch.Tensor, sink_layout: str = "sh",
                               inplace: bool = False) -> torch.Tensor:
    """out rescaled by exp(lse - lse_with_sink) (reference :593-631)."""
    w = _safe_sub(lse, correct_attn_lse_with_sink(lse, sink, sink_layout)).exp().unsqueeze(-1)
    return out.mul_(w) if inplace else (out * w).to(out.dtype)


def correct_attn_out_lse_with_sink(out: torch.Tensor, lse: torch.Tensor, sink: torch.Tensor, sink_layout: str = "sh",
                                   inplace: bool = False) -> tuple[torch.Tensor, torch.Tensor]:
    """(out, lse) merged with the attention sink, one fused kernel (``b200_sink_merge``; reference :634-672)."""
    if not inplace:
        out, lse = out.clone(), lse.clone()
    _lib.sink_merge(out, lse, sink.contiguous(), sink_layout)
    return out, lse
