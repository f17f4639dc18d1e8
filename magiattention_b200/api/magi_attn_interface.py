"""Public API of the CP attention engine — same entry points as the reference
(magi_attention/api/magi_attn_interface.py): magi_attn_varlen_key :160, magi_attn_flex_key :440,
dispatch :887, undispatch :924, calc_attn :1041, get_position_ids :1112, get_most_recent_key :1136,
clear_cache :1157.  Keys are cached per (mask, shapes, group, config) like the reference's
DistAttnRuntimeDict (:73-150) so planning runs once per mask."""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.distributed as dist

from ..common.forward_meta import AttnForwardMeta
from ..common.ranges import AttnRanges
from ..config import DispatchConfig, DistAttnConfig, OverlapConfig  # noqa: F401
from ..dist_attn_runtime_mgr import (DistAttnRuntimeKey, DistAttnRuntimeMgr, init_dist_attn_runtime_key,
                                     init_dist_attn_runtime_mgr)
from .functools import infer_attn_mask_from_cu_seqlens

__all__ = ["magi_attn_varlen_key", "magi_attn_flex_key", "magi_attn_varlen_dispatch", "magi_attn_flex_dispatch", "dispatch",
           "undispatch", "calc_attn", "get_position_ids", "get_most_recent_key", "clear_cache", "get_runtime_mgr",
           "make_flex_key_for_new_mask_after_dispatch", "make_varlen_key_for_new_mask_after_dispatch", "roll", "roll_simple",
           "DistAttnRuntimeDictManager", "dist_attn_runtime_dict_mgr"]

_MAX_CACHE = 100
_mgr_cache: "OrderedDict[DistAttnRuntimeKey, DistAttnRuntimeMgr]" = OrderedDict()
_groups: dict[int, dist.ProcessGroup | None] = {0: None}  # cp_group_id -> group (keys only carry the id)
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

    def keys(self, cp_group: dist.ProcessGroup | None = None):
        return [k for k in _mgr_cache if cp_group is None or k.cp_group_id == id(cp_group)]

    def get_most_recent_key(self, cp_group: dist.ProcessGroup | None = None) -> DistAttnRuntimeKey:
        return get_most_recent_key(cp_group)

    def clear(self, cp_group: dist.ProcessGroup | None = None) -> None:
        clear_cache(cp_group)


dist_attn_runtime_dict_mgr = DistAttnRuntimeDictManager()


def _resolve_group(cp_group_or_mesh) -> dist.ProcessGroup | None:
    if cp_group_or_mesh is None or isinstance(cp_group_or_mesh, dist.ProcessGroup):
        return cp_group_or_mesh
    # DeviceMesh: flat (1-D) meshes only — one NVSwitch domain needs no hierarchy
    return cp_group_or_mesh.get_group()


def magi_attn_flex_key(q_ranges: AttnRanges, k_ranges: AttnRanges, attn_mask_type, total_seqlen_q: int,
                       total_seqlen_k: int, num_heads_q: int, num_heads_kv: int, head_dim: int, pad_size: int,
                       cp_group_or_mesh, dist_attn_config: DistAttnConfig = DistAttnConfig(), is_same_source: bool = True,
                       is_q_permutable: bool = True, is_k_permutable: bool = True,
                       chunk_size: int | None = None) -> DistAttnRuntimeKey:
    """Build (or fetch from the cache) the runtime for a mask given as AttnSlices in GLOBAL coordinates."""
    assert is_same_source and is_q_permutable and is_k_permutable, "only self-attention with permutable q/k is supported"
    group = _resolve_group(cp_group_or_mesh)
    cp_size = 1 if group is None else dist.get_world_size(group)
    cs = chunk_size or dist_attn_config.dispatch_config.chunk_size or DEFAULT_CHUNK_SIZE
    total = total_seqlen_q + pad_size
    assert total % (cs * cp_size) == 0, (
        f"total_seqlen_q + pad_size = {total} must be a multiple of chunk_size * cp_size = {cs * cp_size}; "
        "use compute_pad_size(total_seqlen_q, cp_size, chunk_size)")
    key = init_dist_attn_runtime_key(q_ranges, k_ranges, attn_mask_type, total_seqlen_q, total_seqlen_k, pad_size, cs, group,
                                     num_heads_q, num_heads_kv, head_dim, dist_attn_config)
    _groups[key.cp_group_id] = group
    return _register(key, group)


def _register(key: DistAttnRuntimeKey, group) -> DistAttnRuntimeKey:
    if key in _mgr_cache:
        _mgr_cache.move_to_end(key)
    else:
        _mgr_cache[key] = init_dist_attn_runtime_mgr(key, group)
        while len(_mgr_cache) > _MAX_CACHE:
            _mgr_cache.popitem(last=False)
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


def dispatch(x: torch.Tensor, key: DistAttnRuntimeKey) -> torch.Tensor:
    """Global [total_seqlen, ...] -> this rank's padded, permuted shard."""
    return get_runtime_mgr(key).dispatch_qo(x)


def undispatch(x: torch.Tensor, key: DistAttnRuntimeKey) -> torch.Tensor:
    """This rank's shard -> global [total_seqlen, ...] (all-gather + unpermute + unpad)."""
    return get_runtime_mgr(key).undispatch_qo(x)


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
    """Global position of every local token (for RoPE after dispatch)."""
    mgr = get_runtime_mgr(key)
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    return mgr.get_position_ids(dev)


def get_most_recent_key(cp_group: dist.ProcessGroup | None = None) -> DistAttnRuntimeKey:
    """The most recently created / used key (of ``cp_group`` when given) — reference :1136-1154."""
    for k in reversed(_mgr_cache):
        if cp_group is None or k.cp_group_id == id(cp_group):
            return k
    raise ValueError("no dist attn runtime key has been created yet")


def clear_cache(cp_group: dist.ProcessGroup | None = None) -> None:
    """Drop the cached runtime managers (of ``cp_group`` only when given) — reference :1157-1164."""
    if cp_group is None:
        _mgr_cache.clear()
        return
    for k in [k for k in _mgr_cache if k.cp_group_id == id(cp_group)]:
        del _mgr_cache[k]


def make_flex_key_for_new_mask_after_dispatch(q_ranges: AttnRanges, k_ranges: AttnRanges, attn_mask_type,
                                              key_for_dispatch: DistAttnRuntimeKey, num_heads_q: int | None = None,
                                              num_heads_kv: int | None = None, head_dim: int | None = None,
                                              dist_attn_config: DistAttnConfig | None = None) -> DistAttnRuntimeKey:
    """A key for a SECOND mask (hybrid-attention models) that reuses the dispatch solution of ``key_for_dispatch``: tensors
    dispatched with the old key can be fed to ``calc_attn`` with the new one; kernel slices, transfer tables and overlap
    stages are planned for the new mask (reference :1315-1530).  Only the overlap part of ``dist_attn_config`` applies."""
    mgr = _mgr_cache.get(key_for_dispatch)
    if mgr is None:
        raise ValueError("The dist attn runtime key for dispatch does not exist!")
    kd = key_for_dispatch
    cfg = kd.config if dist_attn_config is None else DistAttnConfig(
        dispatch_config=kd.config.dispatch_config, overlap_config=dist_attn_config.overlap_config,
        high_precision_reduce=dist_attn_config.high_precision_reduce)
    group = _groups[kd.cp_group_id]
    dmask = kd.dispatch_mask or (kd.q_ranges, kd.k_ranges, kd.attn_mask_type)
    key = init_dist_attn_runtime_key(q_ranges, k_ranges, attn_mask_type, kd.total_seqlen_q, kd.total_seqlen_k, kd.pad_size,
                                     kd.chunk_size, group, num_heads_q or kd.num_heads_q, num_heads_kv or kd.num_heads_kv,
                                     head_dim or kd.head_dim, cfg, dispatch_mask=dmask)
    return _register(key, group)


def make_varlen_key_for_new_mask_after_dispatch(cu_seqlens_q: torch.Tensor, cu_seqlens_k: torch.Tensor,
                                                key_for_dispatch: DistAttnRuntimeKey, causal: bool = False,
                                                window_size: tuple[int, int] = (-1, -1),
                                                dist_attn_config: DistAttnConfig | None = None) -> DistAttnRuntimeKey:
    """flash-attn-varlen flavour of :func:`make_flex_key_for_new_mask_after_dispatch` (reference :1167-1312)."""
    qr, kr, ty, _, _ = infer_attn_mask_from_cu_seqlens(cu_seqlens_q, cu_seqlens_k, causal, window_size)
    return make_flex_key_for_new_mask_after_dispatch(qr, kr, ty, key_for_dispatch, dist_attn_config=dist_attn_config)
