"""AttnForwardMeta — second return value of flex_flash_attn_func / calc_attn
(reference: magi_attention/common/forward_meta.py:21-34)."""
from __future__ import annotations

from dataclasses import dataclass

import torch


@dataclass
class AttnForwardMeta:
    lse: torch.Tensor | None
    max_logits: torch.Tensor | None = None
