"""The reference runtime's B200 kernel slot, served by this backend.

On Blackwell the reference's DistAttnRuntime dispatches its stage kernels to ``fa4_fwd`` / ``fa4_bwd``
(magi_attention/functional/fa4.py:125-321, called from functional/dist_attn.py:1230-1324,1328-1460 when
``MAGI_ATTENTION_KERNEL_BACKEND=fa4``).  These two functions keep that slot contract — same names, parameter lists, order,
defaults and return tuples — on top of the sm_100a kernels of this package, so the reference's own runtime can call this
backend without this package's dist_attn.py.  ``attn_arg`` is any object with ``to_ffa_args(is_bwd)`` returning the
``q_ranges`` / ``k_ranges`` / ``attn_type_map`` int32 device tensors (the reference's AttnArg, calc_meta.py:146-159, and this
package's meta.planner.AttnArg both do): slices are consumed directly, no HSTU-function / block-sparse conversion happens.
Returned dtypes follow the slot: ``out`` in q's dtype, ``lse`` fp32 ``[tq, hq]``; ``dq, dk, dv`` in the input dtype.
"""
from __future__ import annotations

import torch

from .flex_flash_attn import _flex_flash_attn_backward, _flex_flash_attn_forward

__all__ = ["fa4_fwd", "fa4_bwd", "is_fa4_installed"]

is_fa4_installed = True  # the slot is always served: the kernels are this package's own library


def _ranges(attn_arg, is_bwd: bool, device) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor | None]:
    a = attn_arg.to_ffa_args(is_bwd) if not hasattr(attn_arg, "_dev") else attn_arg.to_ffa_args(is_bwd, device)
    return a["q_ranges"], a["k_ranges"], a.get("attn_type_map")


@torch.no_grad()
def fa4_fwd(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    sink: torch.Tensor | None,
    attn_arg,
    softmax_scale: float | None = None,
    softcap: float = 0.0,
    sink_layout: str = "sh",
) -> tuple[torch.Tensor, torch.Tensor]:
    if softmax_scale is None:
        softmax_scale = q.shape[-1] ** (-0.5)
    qr, kr, ty = _ranges(attn_arg, False, q.device)
    out, meta = _flex_flash_attn_forward(q, k, v, sink, sink_layout, None, None, qr, kr, ty, softmax_scale, softcap, q.dtype)
    return out, meta.lse


@torch.no_grad()
def fa4_bwd(
    do: torch.Tensor,
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    sink: torch.Tensor | None,
    o: torch.Tensor,
    lse: torch.Tensor,
    attn_arg,
    softmax_scale: float | None = None,
    softcap: float = 0.0,
    sink_layout: str = "sh",
    deterministic: bool = False,
) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor | None]:
    if softmax_scale is None:
        softmax_scale = q.shape[-1] ** (-0.5)
    qr, kr, ty = _ranges(attn_arg, True, q.device)
    dq, dk, dv, dsink = _flex_flash_attn_backward(do, q, k, v, sink, sink_layout, o, lse, None, None, None, None, qr, kr, ty,
                                                  softmax_scale, softcap, dk_type=k.dtype, dv_type=v.dtype,
                                                  deterministic=deterministic)
    return dq.to(q.dtype), dk, dv, dsink
