"""Public API of the CP attention engine — same entry points as the reference
(magi_attention/api/magi_attn_interface.py): magi_attn_varlen_key :160, magi_attn_flex_key :440,
dispatch :887, undispatch :924, calc_attn :1041, get_position_ids :1112, get_most_recent_key :1136,
clear_cache :1157.  Keys are cached per (mask, shapes, group, config) like the reference's
DistAttnRuntimeDict (:73-150) so planning runs once per mask."""
from __future__ import annotations

import dataclasses
import os
import warnings
from collections import OrderedDict

import torch
import torch.distributed as dist

from ..common.enum import AttnMaskType
from ..common.forward_meta import AttnForwardMeta
from ..common.ranges import AttnRanges
from ..config import DispatchConfig, DistAttnConfig, OverlapConfig  # noqa: F401
from ..dist_attn_runtime_mgr import (DistAttnRuntimeKey, DistAttnRuntimeMgr, init_dist_attn_runtime_key,
                                     init_dist_attn_runtime_mgr)
from .functools import apply_padding, compute_pad_size, infer_attn_mask_from_cu_seqlens, pad_at_dim, unpad_at_dim

__all__ = ["magi_attn_varlen_key", "magi_attn_flex_key", "magi_attn_varlen_dispatch", "magi_attn_flex_dispatch", "dispatch",
           "undispatch", "calc_attn", "get_position_ids", "get_most_recent_key", "clear_cache", "get_runtime_mgr",
           "make_flex_key_for_new_mask_after_dispatch", "make_varlen_key_for_new_mask_after_dispatch", "roll", "roll_simple",
           "DistAttnRuntimeDictManager", "dist_attn_runtime_dict_mgr"]

_MAX_CACHE = 100
_mgr_cache: "OrderedDict[DistAttnRuntimeKey, DistAttnRuntimeMgr]" = OrderedDict()
DEFAULT_CHUNK_SIZE = 2048


class DistAttnRuntimeDictManager:
    """Dict-like view of the runtime-manager cache (reference :64-150): ``get`` / ``in`` / ``[]`` / ``keys`` /
    ``get_most_recent_key`` / ``clear``.  One LRU of ``_MAX_CACHE`` entries serves every cp group here."""

    def get(self, key: DistAttnRuntimeKey, default=None):
        return _mgr_cache.get(key, default)

    def __contains__(self, key: DistAttnRuntimeKey) -> bool:
        return key in _mgr_cache

    def __getitem__(self, key: DistAttnRuntimeKey) -> DistAttnRuntimeMgr:
        return _mgr_cache[key]

    def __setitem__(self, key: DistAttnRuntimeKey, value: DistAttnRuntimeMgr) -> None:
        _mgr_cache[key] = value
        _mgr_cache.move_to_end(key)
        while len(_mgr_cache) > _MAX_CACHE:
            _mgr_cache.popitem(last=False)

    def keys(self, cp_group: dist.ProcessGroup | None = None):
        return [k for k in _mgr_cache if cp_group is None or k.cp_group is cp_group]

    def get_most_recent_key(self, cp_group: dist.ProcessGroup | None = None) -> DistAttnRuntimeKey:
        return get_most_recent_key(cp_group)

    def clear(self, cp_group: dist.ProcessGroup | None = None) -> None:
        clear_cache(cp_group)


dist_attn_runtime_dict_mgr = DistAttnRuntimeDictManager()


def _resolve_group(cp_group_or_mesh):
    """-> (cp_group, cp_mesh) like the reference (:620-640); None is accepted for single-process cp=1 use."""
    if cp_group_or_mesh is None or isinstance(cp_group_or_mesh, dist.ProcessGroup):
        return cp_group_or_mesh, None
    mesh = cp_group_or_mesh
    if getattr(mesh, "ndim", 1) != 1:
        raise NotImplementedError("2-D cp_mesh (hierarchical comm) is not supported: one NVSwitch domain is flat")
    return mesh.get_group(), mesh


def _min_chunks_per_rank() -> int:
    """MAGI_ATTENTION_MIN_CHUNKS_PER_RANK (reference env/general.py:263-271), default 8."""
    return int(os.environ.get("MAGI_ATTENTION_MIN_CHUNKS_PER_RANK", "8"))


def magi_attn_flex_key(q_ranges: AttnRanges, k_ranges: AttnRanges, attn_mask_type, total_seqlen_q: int,
                       total_seqlen_k: int, num_heads_q: int, num_heads_kv: int, head_dim: int, pad_size: int,
                       cp_group_or_mesh, dist_attn_config: DistAttnConfig = DistAttnConfig(), is_same_source: bool = True,
                       is_q_permutable: bool = True, is_k_permutable: bool = True,
                       chunk_size: int | None = None) -> DistAttnRuntimeKey:
    """Build (or fetch from the cache) the runtime for a mask given as AttnSlices in GLOBAL coordinates.

    Follows the reference step for step (:573-715): chunk size = the given one (argument or DispatchConfig) capped so that
    every rank gets at least MAGI_ATTENTION_MIN_CHUNKS_PER_RANK chunks, else derived from the sequence length; the pad size
    is computed internally (the ``pad_size`` argument is deprecated there and only checked here); the padding block is
    appended to the mask; the key carries the PADDED total."""
    dispatch_config = dist_attn_config.dispatch_config
    if chunk_size is not None and dispatch_config.chunk_size is not None and dispatch_config.chunk_size != chunk_size:
        raise ValueError(f"Conflicting `chunk_size`: got {chunk_size} from the API parameter and {dispatch_config.chunk_size} "
                         "from `dist_attn_config.dispatch_config`. Please only set it in one place.")
    if chunk_size is None:
        chunk_size = dispatch_config.chunk_size
    assert q_ranges.end <= total_seqlen_q and k_ranges.end <= total_seqlen_k, (
        f"The maximum endpoint in ranges must be less than total_seqlen, but got {q_ranges.end=} when {total_seqlen_q=}, "
        f"and got {k_ranges.end=} when {total_seqlen_k=}")
    group, mesh = _resolve_group(cp_group_or_mesh)
    cp_size = 1 if group is None else dist.get_world_size(group)
    auto = -(-total_seqlen_q // (_min_chunks_per_rank() * cp_size))
    chunk_size = min(auto, chunk_size) if chunk_size is not None else auto
    assert -(-total_seqlen_q // chunk_size) >= cp_size, (
        f"The number of chunks (ceil_div({total_seqlen_q}, {chunk_size})) must be >= cp_size ({cp_size}).")
    computed_pad = compute_pad_size(total_seqlen_q, cp_size, chunk_size)
    if pad_size not in (0, computed_pad):
        warnings.warn(f"The `pad_size` parameter is deprecated: it is computed internally ({computed_pad}, chunk_size "
                      f"{chunk_size}); the given value {pad_size} is ignored.", DeprecationWarning, stacklevel=2)
    pad_size = computed_pad
    n = len(q_ranges)
    types = list(attn_mask_type) if isinstance(attn_mask_type, (list, tuple)) else [attn_mask_type] * n
    if pad_size > 0:
        q_ranges, k_ranges, types = apply_padding(q_ranges, k_ranges, [AttnMaskType(t) if isinstance(t, str) else
                                                                         (t if isinstance(t, AttnMaskType) else
                                                                          AttnMaskType.from_int_type(int(t))) for t in types],
                                                  total_seqlen_q, pad_size)
        total_seqlen_q += pad_size
        total_seqlen_k += pad_size
    key = init_dist_attn_runtime_key(q_ranges, k_ranges, types, total_seqlen_q, total_seqlen_k, num_heads_q, num_heads_kv,
                                     head_dim, pad_size, chunk_size, group, mesh, dist_attn_config)
    if key not in dist_attn_runtime_dict_mgr:
        dist_attn_runtime_dict_mgr[key] = init_dist_attn_runtime_mgr(
            q_ranges, k_ranges, list(key.attn_mask_type), total_seqlen_q, total_seqlen_k, num_heads_q, num_heads_kv, head_dim,
            chunk_size, group, mesh, dist_attn_config, is_same_source, is_q_permutable, is_k_permutable)
    else:
        _mgr_cache.move_to_end(key)
    return key


def magi_attn_varlen_key(cu_seqlens_q: torch.Tensor, cu_seqlens_k: torch.Tensor, num_heads_q: int, num_heads_kv: int,
                         head_dim: int, pad_size: int, cp_group_or_mesh, causal: bool = False,
                         window_size: tuple[int, int] = (-1, -1), dist_attn_config: DistAttnConfig = DistAttnConfig(),
                         chunk_size: int | None = None) -> DistAttnRuntimeKey:
    """flash-attn-varlen-like entry: cu_seqlens (+causal / sliding window) -> key."""
    qr, kr, ty, tq, tk = infer_attn_mask_from_cu_seqlens(cu_seqlens_q, cu_seqlens_k, causal, window_size)
    return magi_attn_flex_key(qr, kr, ty, tq, tk, num_heads_q, num_heads_kv, head_dim, pad_size, cp_group_or_mesh,
                              dist_attn_config, chunk_size=chunk_size)


def get_runtime_mgr(key: DistAttnRuntimeKey) -> DistAttnRuntimeMgr:
    mgr = _mgr_cache.get(key)
    if mgr is None:
        raise ValueError("The dist attn runtime key does not exist (evicted or never created)")
    return mgr


def dispatch(x: torch.Tensor, key: DistAttnRuntimeKey, pad_value: float = 0.0) -> torch.Tensor:
    """Global [total_seqlen, ...] -> this rank's padded, permuted shard (reference :887-921: pad at dim 0 with
    ``pad_value``, then DistAttnRuntimeMgr.dispatch_qo)."""
    return get_runtime_mgr(key).dispatch_qo(pad_at_dim(x, 0, key.pad_size, value=pad_value))


def undispatch(x: torch.Tensor, key: DistAttnRuntimeKey, is_partial_grad: bool = False) -> torch.Tensor:
    """This rank's shard -> global [total_seqlen, ...] (all-gather + unpermute + unpad; reference :924-957).
    ``is_partial_grad``: the gradient flowing back into this op is a per-rank partial sum and is reduced first."""
    return unpad_at_dim(get_runtime_mgr(key).undispatch_qo(x, is_partial_grad=is_partial_grad), 0, key.pad_size)


def roll(x: torch.Tensor, shift: int, dim: int, key: DistAttnRuntimeKey) -> torch.Tensor:
    """Cyclically roll a dispatched local tensor along ``dim`` of the GLOBAL (padded) sequence — multi-token-prediction
    labels; equals undispatch -> torch.roll -> dispatch without materialising the global tensor (reference :960-1001)."""
    return get_runtime_mgr(key).roll(x, shift, dim)


def roll_simple(x: torch.Tensor, shift: int, dim: int, key: DistAttnRuntimeKey) -> torch.Tensor:
    """:func:`roll` with plain isend / irecv instead of one grouped batch (reference :1004-1038)."""
    return get_runtime_mgr(key).roll_simple(x, shift, dim)


def magi_attn_flex_dispatch(x: torch.Tensor, *args, **kwargs) -> tuple[torch.Tensor, DistAttnRuntimeKey]:
    key = magi_attn_flex_key(*args, **kwargs)
    return dispatch(x, key), key


def magi_attn_varlen_dispatch(x: torch.Tensor, *args, **kwargs) -> tuple[torch.Tensor, DistAttnRuntimeKey]:
    key = magi_attn_varlen_key(*args, **kwargs)
    return dispatch(x, key), key


def calc_attn(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, key: DistAttnRuntimeKey, sink: torch.Tensor | None = None,
              softmax_scale: float | None = None, softcap: float = 0.0,
              return_max_logits: bool = False) -> tuple[torch.Tensor, AttnForwardMeta]:
    """Distributed attention on local shards [shard_seqlen, heads, head_dim] (reference :1041-1109)."""
    return get_runtime_mgr(key).calc_attn(q, k, v, sink=sink, softmax_scale=softmax_scale, softcap=softcap,
                                          return_max_logits=return_max_logits)


def get_position_ids(key: DistAttnRuntimeKey) -> torch.Tensor:
    """Global position of every local token (for RoPE after dispatch) — reference :1112-1133."""
    mgr = get_runtime_mgr(key)
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    return mgr.get_position_ids().to(dev)


def get_most_recent_key(cp_group: dist.ProcessGroup | None = None) -> DistAttnRuntimeKey:
    """The most recently created / used key (of ``cp_group`` when given) — reference :1136-1154."""
    for k in reversed(_mgr_cache):
        if cp_group is None or k.cp_group is cp_group:
            return k
    raise ValueError("no dist attn runtime key has been created yet")


def clear_cache(cp_group: dist.ProcessGroup | None = None) -> None:
    """Drop the cached runtime managers (of ``cp_group`` only when given) — reference :1157-1164."""
    if cp_group is None:
        _mgr_cache.clear()
        return
    for k in [k for k in _mgr_cache if k.cp_group is cp_group]:
        del _mgr_cache[k]


def make_flex_key_for_new_mask_after_dispatch(q_ranges: AttnRanges, k_ranges: AttnRanges, attn_mask_type,
                                              key_for_dispatch: DistAttnRuntimeKey, num_heads_q: int | None = None,
                                              num_heads_kv: int | None = None, head_dim: int | None = None,
                                              dist_attn_config: DistAttnConfig | None = None) -> DistAttnRuntimeKey:
    """A key for a SECOND mask (hybrid-attention models) that reuses the dispatch solution of ``key_for_dispatch``: tensors
    dispatched with the old key can be fed to ``calc_attn`` with the new one; kernel slices, transfer tables and overlap
    stages are planned for the new mask (reference :1315-1530: passes ref_dispatch_meta_q/k to init_dist_attn_runtime_mgr).
    Only the overlap part of ``dist_attn_config`` applies."""
    mgr = _mgr_cache.get(key_for_dispatch)
    if mgr is None:
        raise ValueError("The dist attn runtime key for dispatch does not exist!")
    kd = key_for_dispatch
    cfg = kd.dist_attn_config if dist_attn_config is None else DistAttnConfig(
        dispatch_config=kd.dist_attn_config.dispatch_config, overlap_config=dist_attn_config.overlap_config,
        high_precision_reduce=dist_attn_config.high_precision_reduce)
    n = len(q_ranges)
    types = list(attn_mask_type) if isinstance(attn_mask_type, (list, tuple)) else [attn_mask_type] * n
    types = [AttnMaskType(t) if isinstance(t, str) else (t if isinstance(t, AttnMaskType) else AttnMaskType.from_int_type(int(t)))
             for t in types]
    if kd.pad_size > 0:
        q_ranges, k_ranges, types = apply_padding(q_ranges, k_ranges, types, kd.total_seqlen_q - kd.pad_size, kd.pad_size)
    hq, hk, hd = num_heads_q or kd.num_heads_q, num_heads_kv or kd.num_heads_kv, head_dim or kd.head_dim
    dmask = kd.dispatch_mask or (kd.q_ranges, kd.k_ranges, kd.attn_mask_type)
    key = dataclasses.replace(
        init_dist_attn_runtime_key(q_ranges, k_ranges, types, kd.total_seqlen_q, kd.total_seqlen_k, hq, hk, hd, kd.pad_size,
                                   kd.chunk_size, kd.cp_group, kd.cp_mesh, cfg), dispatch_mask=dmask)
    if key not in dist_attn_runtime_dict_mgr:
        dist_attn_runtime_dict_mgr[key] = init_dist_attn_runtime_mgr(
            q_ranges, k_ranges, types, kd.total_seqlen_q, kd.total_seqlen_k, hq, hk, hd, kd.chunk_size, kd.cp_group, kd.cp_mesh,
            cfg, ref_dispatch_meta_q=mgr.dispatch_meta_q, ref_dispatch_meta_k=mgr.dispatch_meta_k)
    return key


def make_varlen_key_for_new_mask_after_dispatch(cu_seqlens_q: torch.Tensor, cu_seqlens_k: torch.Tensor,
                                                key_for_dispatch: DistAttnRuntimeKey, causal: bool = False,
                                                window_size: tuple[int, int] = (-1, -1),
                                                dist_attn_config: DistAttnConfig | None = None) -> DistAttnRuntimeKey:
    """flash-attn-varlen flavour of :func:`make_flex_key_for_new_mask_after_dispatch` (reference :1167-1312)."""
    qr, kr, ty, _, _ = infer_attn_mask_from_cu_seqlens(cu_seqlens_q, cu_seqlens_k, causal, window_size)
    return make_flex_key_for_new_mask_after_dispatch(qr, kr, ty, key_for_dispatch, dist_attn_config=dist_attn_config)
