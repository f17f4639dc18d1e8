"""API helpers (reference: magi_attention/api/functools.py): pad size, batch squashing, and
cu_seqlens / sliding-window -> AttnSlice inference."""
from __future__ import annotations

import torch

from ..common.enum import AttnMaskType
from ..common.range import AttnRange
from ..common.ranges import AttnRanges
from ..meta.planner import _restrict

__all__ = ["compute_pad_size", "squash_batch_dim", "infer_varlen_mask_from_batch", "infer_attn_mask_from_sliding_window",
           "infer_attn_mask_from_cu_seqlens", "pad_at_dim", "unpad_at_dim", "apply_padding"]


def compute_pad_size(total_seqlen_q: int, cp_size: int, chunk_size: int) -> int:
    """Tokens to append so the sequence is a multiple of chunk_size * cp_size (reference :27-51)."""
    block = chunk_size * cp_size
    rem = total_seqlen_q % block
    return 0 if rem == 0 else block - rem


def squash_batch_dim(x: torch.Tensor) -> torch.Tensor:
    """[b, s, ...] -> [b*s, ...] (reference :54-65)."""
    return x.reshape(x.shape[0] * x.shape[1], *x.shape[2:])


def infer_varlen_mask_from_batch(batch_size: int, seq_len: int, device: str | None = None) -> tuple[torch.Tensor, torch.Tensor]:
    """cu_seqlens_q/k (int32 ``[batch_size + 1]``) of a [b, s] batch flattened to one packed sequence (reference :68-91;
    ``device`` defaults to "cuda" like the reference when a GPU is present)."""
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    cu = torch.arange(0, batch_size + 1, dtype=torch.int32, device=device) * seq_len
    return cu, cu


def apply_padding(q_ranges: AttnRanges, k_ranges: AttnRanges, attn_mask_type: list, total_seqlen: int,
                  pad_size: int) -> tuple[AttnRanges, AttnRanges, list]:
    """Append the padding rows ``[total_seqlen, total_seqlen + pad_size)`` as a q range with an empty k range and a FULL type
    — a slice that attends to nothing, so padded rows come out as 0 / -inf (reference :142-177).  The key builders here do
    not need it (uncovered rows already behave that way); it exists for callers that build slice lists themselves."""
    qr = AttnRanges.from_ranges(q_ranges.to_naive_ranges())
    kr = AttnRanges.from_ranges(k_ranges.to_naive_ranges())
    ty = list(attn_mask_type)
    qr.append(AttnRange(total_seqlen, total_seqlen + pad_size))
    kr.append(AttnRange(0, 0))
    ty.append(AttnMaskType.FULL)
    return qr, kr, ty


def pad_at_dim(x: torch.Tensor, dim: int, pad_size: int, value: float = 0.0, side: str = "right") -> torch.Tensor:
    if pad_size == 0:
        return x
    shape = list(x.shape)
    shape[dim] = pad_size
    pad = x.new_full(shape, value)
    return torch.cat([x, pad] if side == "right" else [pad, x], dim=dim)


def unpad_at_dim(x: torch.Tensor, dim: int, pad_size: int) -> torch.Tensor:
    return x if pad_size == 0 else x.narrow(dim, 0, x.shape[dim] - pad_size)


def infer_attn_mask_from_sliding_window(q_range: AttnRange, k_range: AttnRange, window_size: tuple[int, int],
                                        ) -> tuple[AttnRanges, AttnRanges, list[AttnMaskType]]:
    """One (bottom-right aligned) sliding-window block -> AttnSlices (reference :180-332).  Row i attends
    column j iff  i + d - wl <= j <= i + d + wr  with d = sk - sq (flash-attn convention; -1 = unbounded).
    The band is a bi-causal slice of a virtually widened key range; cutting it back to ``k_range`` with the
    planner's exact band decomposition yields CAUSAL / BICAUSAL|FULL / INVCAUSAL pieces."""
    wl, wr = window_size
    q0, q1, k0, k1 = q_range.start, q_range.end, k_range.start, k_range.end
    if q1 - q0 > k1 - k0:
        q0 = q1 - (k1 - k0)  # rows above the bottom-right aligned square see no key (reference :209-214)
    sq, sk = q1 - q0, k1 - k0
    if wl < 0 or wl >= sk - 1:
        wl = -1
    if wr < 0 or wr >= sk - 1:
        wr = -1
    t = (1 if wr >= 0 else 0) | (2 if wl >= 0 else 0)
    vk0 = k0 + (sk - sq) - wl if wl >= 0 else k0
    vk1 = k1 + wr if wr >= 0 else k1
    if t == 0:
        pieces = [(q0, q1, k0, k1, 0)] if sq > 0 and sk > 0 else []
    else:
        pieces = _restrict((q0, q1, vk0, vk1, t), q0, q1, k0, k1)
    qr, kr, ty = AttnRanges(), AttnRanges(), []
    for (a, b, c, e, tt) in pieces:
        qr.append(AttnRange(a, b))
        kr.append(AttnRange(c, e))
        ty.append(AttnMaskType.from_int_type(tt))
    return qr, kr, ty


def infer_attn_mask_from_cu_seqlens(cu_seqlens_q, cu_seqlens_k, causal: bool = False,
                                    window_size: tuple[int, int] = (-1, -1),
                                    ) -> tuple[AttnRanges, AttnRanges, list[AttnMaskType], int, int]:
    """Packed varlen batch -> (q_ranges, k_ranges, mask types, total_q, total_k) (reference :335-430)."""
    cq = [int(x) for x in (cu_seqlens_q.tolist() if isinstance(cu_seqlens_q, torch.Tensor) else cu_seqlens_q)]
    ck = [int(x) for x in (cu_seqlens_k.tolist() if isinstance(cu_seqlens_k, torch.Tensor) else cu_seqlens_k)]
    assert len(cq) == len(ck)
    qr, kr, ty = AttnRanges(), AttnRanges(), []
    for i in range(len(cq) - 1):
        q, k = AttnRange(cq[i], cq[i + 1]), AttnRange(ck[i], ck[i + 1])
        if q.is_empty() or k.is_empty():
            continue
        if window_size != (-1, -1):
            assert not causal, "causal and window_size are mutually exclusive (use window_size=(w, 0))"
            a, b, c = infer_attn_mask_from_sliding_window(q, k, window_size)
            qr.extend(a); kr.extend(b); ty.extend(c)
        else:
            qr.append(q); kr.append(k)
            ty.append(AttnMaskType.CAUSAL if causal else AttnMaskType.FULL)
    return qr, kr, ty, cq[-1], ck[-1]
