"""Slice-metadata fixtures shared by the oracle, golden and GPU parity tests.

The shapes follow the reference's own kernel test masks
(/root/reference/tests/test_attn/test_flex_flash_attn.py:1341-1640: full, varlen_full,
block_causal, varlen_block_causal, sparse_attn (+disjoint / same-k variants),
deterministic_sample) plus the four mask types x three aspect ratios pictured in
functional/flex_flash_attn.py:1252-1340 and the sliding-window decomposition of
api/functools.py:277-329.  Types: 0 full, 1 causal, 2 inv-causal,
3 bi-causal.
"""
from __future__ import annotations


def _scale(ranges, s):
    return [[a // s, b // s] for a, b in ranges]


def _drop_empty(cfg):
    keep = [i for i, (q, k) in enumerate(zip(cfg["q_ranges"], cfg["k_ranges"])) if q[1] > q[0] and k[1] > k[0]]
    for key in ("q_ranges", "k_ranges", "types"):
        cfg[key] = [cfg[key][i] for i in keep]
    return cfg


_VARLEN_1K = [[0, 366], [366, 391], [391, 471], [471, 835], [835, 984], [984, 1005], [1005, 1017], [1017, 1020],
              [1020, 1023], [1023, 1024]]
_BLK_Q_2K = [[0, 256], [256, 512], [512, 1024], [1024, 1280], [1280, 1536], [1536, 1792], [1792, 2048]]
_SPARSE_Q = [[0, 256], [0, 256], [0, 256], [256, 512], [256, 512], [512, 1024], [1024, 1280], [1280, 1536],
             [1280, 1536], [1280, 1536], [1280, 1536], [1280, 1536], [1536, 1792], [1792, 2048]]
_SPARSE_K = [[0, 256], [512, 768], [1011, 1123], [0, 512], [777, 888], [0, 1024], [1024, 1280], [0, 128],
             [555, 556], [777, 982], [1024, 1536], [1689, 1898], [1024, 1792], [1024, 2048]]


def mask_configs(scale: int = 1) -> list[dict]:
    """All configs; `scale` divides every coordinate (golden fixtures use scale=4)."""
    s = scale
    cfgs = [
        dict(name="full", seqlen=2048 // s, q_ranges=[[0, 2048 // s]], k_ranges=[[0, 2048 // s]], types=[0]),
        dict(name="causal", seqlen=2048 // s, q_ranges=[[0, 2048 // s]], k_ranges=[[0, 2048 // s]], types=[1]),
        dict(name="varlen_full_1k", seqlen=1024 // s, q_ranges=_scale(_VARLEN_1K, s), k_ranges=_scale(_VARLEN_1K, s),
             types=[0] * 10),
        dict(name="varlen_causal_1k", seqlen=1024 // s, q_ranges=_scale(_VARLEN_1K, s),
             k_ranges=_scale(_VARLEN_1K, s), types=[1] * 10),
        dict(name="block_causal_2k", seqlen=2048 // s, q_ranges=_scale(_BLK_Q_2K, s),
             k_ranges=_scale([[0, 256], [0, 512], [0, 1024], [0, 1280], [0, 1536], [0, 1792], [0, 2048]], s),
             types=[0] * 7),
        dict(name="varlen_block_causal_2k", seqlen=2048 // s, q_ranges=_scale(_BLK_Q_2K, s),
             k_ranges=_scale([[0, 256], [0, 512], [0, 1024], [1024, 1280], [1024, 1536], [1024, 1792], [1024, 2048]], s),
             types=[0] * 7),
        dict(name="sparse_attn_2k", seqlen=2048 // s, q_ranges=_scale(_SPARSE_Q, s), k_ranges=_scale(_SPARSE_K, s),
             types=[0] * 14),
        dict(name="sparse_attn_2k_disjoint", seqlen=2048 // s,
             q_ranges=_scale(_SPARSE_Q[:5] + _SPARSE_Q[6:], s), k_ranges=_scale(_SPARSE_K[:5] + _SPARSE_K[6:], s),
             types=[0] * 13),
        # every mask type at sq<sk, sq==sk, sq>sk, packed along q with private k ranges
        dict(name="all_types_aspects", seqlen=2304 // s,
             q_ranges=_scale([[0, 128], [128, 384], [384, 768], [768, 896], [896, 1152], [1152, 1536], [1536, 1664],
                              [1664, 1920], [1920, 2304]], s),
             k_ranges=_scale([[0, 384], [384, 640], [640, 768], [768, 1152], [1152, 1408], [1408, 1536],
                              [1536, 1920], [1920, 2176], [2176, 2304]], s),
             types=[1, 1, 1, 2, 2, 2, 3, 3, 3]),
        # sliding-window-causal decomposition: causal head + bi-causal band per document
        dict(name="sliding_window_causal", seqlen=2048 // s,
             q_ranges=_scale([[0, 256], [256, 752], [752, 880], [880, 1136], [1136, 2048]], s),
             k_ranges=_scale([[0, 256], [0, 752], [752, 880], [880, 1136], [880, 2048]], s),
             types=[1, 3, 1, 1, 3]),
        # 2-D grid of small blocks alternating full / causal (reference `deterministic_sample`)
        dict(name="grid_full_causal", seqlen=1200 // s,
             q_ranges=[[i * (200 // s), (i + 1) * (200 // s)] for i in range(6) for _ in range(6)],
             k_ranges=[[j * (200 // s), (j + 1) * (200 // s)] for _ in range(6) for j in range(6)],
             types=[0, 1] * 18),
        # uncovered q rows (O = 0, LSE = -inf)
        dict(name="uncovered_rows", seqlen=1024 // s, q_ranges=_scale([[0, 200], [456, 800]], s),
             k_ranges=_scale([[100, 612], [0, 300]], s), types=[1, 0]),
    ]
    return [_drop_empty(c) for c in cfgs]
