"""Synthetic workloads of the five BASELINE.json configs, as AttnSlice lists in GLOBAL coordinates.

Every generator is deterministic (seed 42) and pure Python, so `bench.py`, the parity tests and the reference arm all see
byte-identical masks.  `area` = unmasked (q, k) pairs; FLOPs_fwd = 4 * area * nhq * d, bwd = 2.5x
(reference: exps/dist_attn/benchmark/metric.py:30-54).
"""
from __future__ import annotations

import random

import numpy as np

D = 128

# Document-length histogram of the reference's default benchmark dataset
# (exps/dist_attn/benchmark/datasets/default/doc_length_distribution.csv: interval -> number of documents); sampled the way
# the reference samples it (exps/dist_attn/benchmark/utils.py:348-400 generate_seqlens: pick an interval by weight, then a
# uniform length inside it, clip to the remaining tokens).
DOC_LENGTH_HISTOGRAM = {
    (0, 10): 30254, (11, 50): 391188, (51, 100): 699596, (101, 200): 1377757, (201, 500): 3794735, (501, 1024): 3316785,
    (1025, 2048): 2106566, (2049, 4096): 988525, (4097, 8192): 410002, (8193, 16384): 173574, (16385, 32768): 115007,
    (32769, 65536): 18428, (65537, 131072): 22513, (131073, 196608): 18284,
}


def slices_area(q_ranges, k_ranges, types) -> int:
    """Unmasked (q, k) pairs of a slice list (area-disjoint slices)."""
    total = 0
    for (q0, q1), (k0, k1), t in zip(q_ranges, k_ranges, types):
        sq, sk = q1 - q0, k1 - k0
        if sq <= 0 or sk <= 0:
            continue
        if t == 0:
            total += sq * sk
            continue
        i = np.arange(sq, dtype=np.int64)
        lo = i if t & 2 else np.zeros_like(i)                    # inv-causal edge: j >= i
        hi = i + (sk - sq) + 1 if t & 1 else np.full_like(i, sk)  # causal edge:     j <= i + sk - sq
        total += int(np.clip(np.minimum(hi, sk) - np.maximum(lo, 0), 0, None).sum())
    return total


def sample_doc_lengths(total: int, cap: int, seed: int = 42) -> list[int]:
    rng = random.Random(seed)
    items = list(DOC_LENGTH_HISTOGRAM.items())
    intervals = [k for k, _ in items]
    weights = [v for _, v in items]
    lens, cur = [], 0
    while cur < total:
        a, b = rng.choices(intervals, weights=weights, k=1)[0]
        n = max(1, min(rng.randint(a, b), cap, total - cur))
        lens.append(n)
        cur += n
    return lens


def sliding_window_causal_slices(a: int, b: int, w: int):
    """Document [a, b) with a causal sliding window of w keys to the left (flash-attn window_size=(w, 0)): row i sees
    keys [max(a, i - w), i].  Same decomposition the reference's infer_attn_mask_from_sliding_window emits
    (api/functools.py:309-329, pinned in tests/golden/functools_masks.json): a CAUSAL head over the first w+1 rows and a
    BICAUSAL band for the rest."""
    n = b - a
    if n <= w + 1:
        return [((a, b), (a, b), 1)]
    h = w + 1
    return [((a, a + h), (a, a + h), 1), ((a + h, b), (a + 1, b), 3)]


def make_workload(config: str, seqlen: int | None = None, heads: int | None = None) -> dict:
    """-> dict(name, seq_len, num_heads_q, num_heads_kv, head_dim, q_ranges, k_ranges, types, area, mask, fwd_only, cp)."""
    c = config.upper()
    fwd_only = False
    if c == "C1":
        S, hq, hk, cp = seqlen or 4096, heads or 16, heads or 16, 1
        qr, kr, ty, mask = [(0, S)], [(0, S)], [1], "causal"
        fwd_only = True
    elif c in ("C3", "C4"):
        S = seqlen or (65536 if c == "C3" else 262144)
        hq = hk = heads or 32
        cp = 2 if c == "C3" else 8
        qr, kr, ty, mask = [(0, S)], [(0, S)], [1], "causal"
    elif c == "C2":
        S, hq, hk, cp = seqlen or 65536, heads or 32, heads or 32, 1
        lens = sample_doc_lengths(S, S // 4)
        qr, kr, ty, off = [], [], [], 0
        for i, n in enumerate(lens):
            if i % 2 == 0:
                qr.append((off, off + n)); kr.append((off, off + n)); ty.append(1)
            else:
                for q_, k_, t_ in sliding_window_causal_slices(off, off + n, 1024):
                    qr.append(q_); kr.append(k_); ty.append(t_)
            off += n
        mask = f"packed varlen: {len(lens)} docs (cap {S // 4}), even docs causal, odd docs sliding-window(1024)-causal"
    elif c == "C5":
        S, hq, hk, cp = seqlen or 1048576, heads or 64, 8, 8
        rng = random.Random(42)
        cuts = [0]
        while cuts[-1] < S:
            cuts.append(min(S, cuts[-1] + 512 * rng.randint(1, 8)))
        blocks = list(zip(cuts[:-1], cuts[1:]))
        qr, kr = [], []
        for i, (a, b) in enumerate(blocks):
            for j in sorted({0, i - 2, i - 1, i}):
                if j >= 0:
                    qr.append((a, b)); kr.append(blocks[j])
        ty = [0] * len(qr)
        mask = (f"block-sparse video-gen: {len(blocks)} variable blocks of 512..4096 tokens, each q block attends itself, the two "
                f"previous blocks and the first (global) block as FULL rectangles")
    else:
        raise ValueError(f"unknown config {config}")
    area = slices_area(qr, kr, ty)
    return dict(name=c, seq_len=S, num_heads_q=hq, num_heads_kv=hk, head_dim=D, q_ranges=[list(x) for x in qr],
                k_ranges=[list(x) for x in kr], types=list(ty), area=area, mask=mask, fwd_only=fwd_only, cp=cp,
                num_slices=len(ty))
