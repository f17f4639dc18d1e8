"""Host-side CP planner: global mask + cp_size -> per-rank dispatch, per-stage kernel slices and KV
transfer tables.

Takes the place of the reference's planning layer for the masks the hot path serves
(magi_attention/meta/_make_dispatch_meta.py:56-448 min-heap chunk dispatch;
meta/solver/dist_attn_solver.py:297-1920 host/remote slice maker + transfer table + overlap stages;
meta/collection/{dispatch_meta,calc_meta,comm_meta}.py).  It is deliberately a small, closed-form
design instead of a port of the 14 kLoC solver:

  * every rank computes the WHOLE plan deterministically from the (identical) global mask, so planning
    needs no communication at all (the reference all-gathers solver objects);
  * a slice restricted to a row span x column span is re-expressed exactly as <= 4 row bands, each again
    one of the four slice types (see ``_restrict``), so arbitrary (q_range, k_range, type) masks work;
  * remote KV is requested zero-redundantly (only rows some local query row can see), grouped into
    per-owner contiguous segments and split into ``overlap_degree`` stages of equal attention area.

Everything here runs once per mask (cached by key in dist_attn_runtime_mgr) and is off the timed path.
"""
from __future__ import annotations

import heapq
from dataclasses import dataclass, field

import numpy as np
import torch

__all__ = ["DispatchMeta", "AttnArg", "GroupCollectiveArg", "CalcMeta", "CommMeta", "make_dispatch_meta",
           "make_attn_meta", "slice_area"]


# ------------------------------------------------------------------------------------------------
# slice geometry (global coordinates) — same bounds as csrc/ffa_common.cuh
# ------------------------------------------------------------------------------------------------
def _lo(s, r):
    q0, q1, k0, k1, t = s
    return k0 + (r - q0) if t & 2 else k0


def _hi(s, r):
    q0, q1, k0, k1, t = s
    return k1 - (q1 - 1 - r) if t & 1 else k1


def slice_area(s, a: int | None = None, b: int | None = None, c: int | None = None, e: int | None = None) -> int:
    """Unmasked (row, col) pairs of slice ``s`` inside rows [a,b) x cols [c,e) (defaults: whole slice)."""
    q0, q1, k0, k1, t = s
    a = q0 if a is None else max(a, q0)
    b = q1 if b is None else min(b, q1)
    c = k0 if c is None else max(c, k0)
    e = k1 if e is None else min(e, k1)
    if a >= b or c >= e:
        return 0
    r = np.arange(a, b, dtype=np.int64)
    lo = np.maximum((k0 + (r - q0)) if t & 2 else np.full_like(r, k0), c)
    hi = np.minimum((k1 - (q1 - 1 - r)) if t & 1 else np.full_like(r, k1), e)
    return int(np.clip(hi - lo, 0, None).sum())


def _restrict(s, a: int, b: int, c: int, e: int) -> list[tuple[int, int, int, int, int]]:
    """Exact re-expression of slice ``s`` restricted to rows [a,b) x cols [c,e) as row bands that are
    themselves slices (q0,q1,k0,k1,type).  Within a band the causal bound is either clamped by ``e``
    everywhere or nowhere, and the inv-causal bound by ``c`` likewise, so each band is FULL / CAUSAL /
    INVCAUSAL / BICAUSAL with properly aligned k range."""
    q0, q1, k0, k1, t = s
    a, b, c, e = max(a, q0), min(b, q1), max(c, k0), min(e, k1)
    if a >= b or c >= e:
        return []
    # row thresholds (bounds are non-decreasing in r with slope 1)
    # upper bound hi(r): empty while hi(r) <= c ; active (hi(r) <= e) ; clamped (hi(r) > e)
    if t & 1:
        r_up_nonempty = q1 - 1 - (k1 - c) + 1  # first r with hi(r) > c  <=> r > q1-1-(k1-c)
        r_up_clamped = q1 - (k1 - e)           # first r with hi(r) > e
    else:
        r_up_nonempty, r_up_clamped = a, a
    # lower bound lo(r): clamped (lo(r) <= c) ; active (c < lo(r) < e) ; empty once lo(r) >= e
    if t & 2:
        r_lo_active = q0 + (c - k0) + 1        # first r with lo(r) > c
        r_lo_empty = q0 + (e - k0)             # first r with lo(r) >= e
    else:
        r_lo_active, r_lo_empty = b, b
    lo_row, hi_row = max(a, r_up_nonempty), min(b, r_lo_empty)
    if lo_row >= hi_row:
        return []
    cuts = sorted({lo_row, hi_row, *(x for x in (r_up_clamped, r_lo_active) if lo_row < x < hi_row)})
    out = []
    for x, y in zip(cuts[:-1], cuts[1:]):
        up_active = bool(t & 1) and x < r_up_clamped      # whole band has hi(r) <= e
        lo_active = bool(t & 2) and x >= r_lo_active      # whole band has lo(r) > c
        kk0 = _lo(s, x) if lo_active else c
        kk1 = _hi(s, y - 1) if up_active else e
        if kk1 <= kk0:
            continue
        out.append((x, y, kk0, kk1, (1 if up_active else 0) | (2 if lo_active else 0)))
    return out


# ------------------------------------------------------------------------------------------------
# dispatch
# ------------------------------------------------------------------------------------------------
@dataclass
class DispatchMeta:
    """Which chunks each rank owns (reference: meta/collection/dispatch_meta.py:24)."""

    total_seqlen: int
    chunk_size: int
    cp_size: int
    partitions: list[list[int]]  # per rank: sorted chunk ids

    @property
    def num_chunks(self) -> int:
        return self.total_seqlen // self.chunk_size

    @property
    def shard_seqlen(self) -> int:
        return self.total_seqlen // self.cp_size

    def position_ids(self, rank: int) -> torch.Tensor:
        """Global token index of every local row (reference: dispatch_meta.py:92-111)."""
        cs = self.chunk_size
        return torch.cat([torch.arange(c * cs, (c + 1) * cs, dtype=torch.int64) for c in self.partitions[rank]])

    def host_spans(self, rank: int) -> list[tuple[int, int, int]]:
        """Maximal globally-contiguous runs of the rank's chunks: (global_start, global_end, local_offset)."""
        cs = self.chunk_size
        spans: list[tuple[int, int, int]] = []
        for i, c in enumerate(self.partitions[rank]):
            g0, g1, l0 = c * cs, (c + 1) * cs, i * cs
            if spans and spans[-1][1] == g0:
                s = spans[-1]
                spans[-1] = (s[0], g1, s[2])
            else:
                spans.append((g0, g1, l0))
        return spans


def make_dispatch_meta(slices: list[tuple[int, int, int, int, int]], total_seqlen: int, chunk_size: int,
                       cp_size: int, alg: str = "min_heap") -> DispatchMeta:
    """Chunk -> rank assignment with an equal chunk count per rank.

    ``min_heap`` (reference default, meta/_make_dispatch_meta.py:221, meta/solver/dispatch_solver.py:894-1000): chunks in
    decreasing mask area go to the least-loaded rank that still has room.  ``sequential`` (:1178-1212): rank r owns the
    r-th contiguous block of chunks (no balancing; what ring / Ulysses style CP would see).  ``sorted_sequential``
    (SortedSequentialSelectAlg): area-sorted chunks dealt out in snake order 0..cp-1, cp-1..0, ..."""
    assert total_seqlen % (chunk_size * cp_size) == 0, "total_seqlen must be a multiple of chunk_size * cp_size"
    assert alg in ("min_heap", "sequential", "sorted_sequential"), f"unknown dispatch algorithm {alg!r}"
    n_chunks = total_seqlen // chunk_size
    per_rank = n_chunks // cp_size
    areas = np.zeros(n_chunks, dtype=np.int64)
    for s in slices:
        q0, q1 = s[0], s[1]
        for ch in range(q0 // chunk_size, (max(q1, q0 + 1) - 1) // chunk_size + 1):
            areas[ch] += slice_area(s, ch * chunk_size, (ch + 1) * chunk_size)
    if alg == "sequential":
        return DispatchMeta(total_seqlen, chunk_size, cp_size,
                            [list(range(r * per_rank, (r + 1) * per_rank)) for r in range(cp_size)])
    order = sorted(range(n_chunks), key=lambda i: (-int(areas[i]), i))
    if alg == "sorted_sequential":
        snake: list[list[int]] = [[] for _ in range(cp_size)]
        for i, ch in enumerate(order):
            rnd, pos = divmod(i, cp_size)
            snake[pos if rnd % 2 == 0 else cp_size - 1 - pos].append(ch)
        return DispatchMeta(total_seqlen, chunk_size, cp_size, [sorted(p) for p in snake])
    heap = [(0, r) for r in range(cp_size)]
    heapq.heapify(heap)
    parts: list[list[int]] = [[] for _ in range(cp_size)]
    for ch in order:
        popped = []
        while True:
            load, r = heapq.heappop(heap)
            if len(parts[r]) < per_rank:
                break
            popped.append((load, r))
        parts[r].append(ch)
        heapq.heappush(heap, (load + int(areas[ch]), r))
        for it in popped:
            heapq.heappush(heap, it)
    return DispatchMeta(total_seqlen, chunk_size, cp_size, [sorted(p) for p in parts])


# ------------------------------------------------------------------------------------------------
# attention / communication metadata
# ------------------------------------------------------------------------------------------------
@dataclass
class AttnArg:
    """Slices of one kernel launch in LOCAL coordinates (reference: meta/collection/calc_meta.py:62)."""

    q_ranges: list[tuple[int, int]] = field(default_factory=list)
    k_ranges: list[tuple[int, int]] = field(default_factory=list)
    attn_type_map: list[int] = field(default_factory=list)
    total_area: int = 0
    _dev: dict = field(default_factory=dict, repr=False)

    @property
    def can_skip(self) -> bool:
        return len(self.q_ranges) == 0

    def to_ffa_args(self, is_bwd=False, device=None) -> dict[str, torch.Tensor]:
        """int32 device tensors (cached) — reference AttnArg.to_ffa_args(is_bwd), calc_meta.py:146-159,210-223.
        Forward and backward use the same slice list here (the kernels build their own tile->slice plans), so ``is_bwd``
        only exists for signature parity; a torch.device passed first is taken as ``device`` (default: current CUDA device)."""
        if isinstance(is_bwd, (torch.device, str)):
            device = is_bwd
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        key = str(device)
        if key not in self._dev:
            mk = lambda x, shape: torch.tensor(x, dtype=torch.int32, device=device).reshape(shape)
            self._dev[key] = dict(q_ranges=mk(self.q_ranges, (-1, 2)), k_ranges=mk(self.k_ranges, (-1, 2)),
                                  attn_type_map=mk(self.attn_type_map, (-1,)))
        return self._dev[key]


@dataclass
class GroupCollectiveArg:
    """Arguments of one group_cast (and, transposed, the matching group_reduce) — reference:
    meta/collection/comm_meta.py:41-155."""

    input_split_size_list: list[int]
    output_split_size_list: list[int]
    dst_indices_list: list[list[int]]
    src_index_list: list[int]

    def to_group_cast_args(self) -> dict:
        return dict(input_split_sizes=self.input_split_size_list, output_split_sizes=self.output_split_size_list,
                    dst_indices=self.dst_indices_list, src_index=self.src_index_list)

    def to_group_reduce_args(self) -> dict:
        return dict(input_split_sizes=self.output_split_size_list, output_split_sizes=self.input_split_size_list,
                    dst_index=self.src_index_list, src_indices=self.dst_indices_list)


@dataclass
class CalcMeta:
    local_attn_arg: AttnArg
    remote_attn_args_list: list[AttnArg]
    seqlen_q_shard: int
    seqlen_k_local: int
    seqlen_k_remote_per_stage: list[int]

    @property
    def overlap_degree(self) -> int:
        return len(self.remote_attn_args_list)


@dataclass
class CommMeta:
    kv_group_collective_args_list: list[GroupCollectiveArg]
    num_remote_kv_tokens_per_stage: list[int]


def make_attn_meta(slices: list[tuple[int, int, int, int, int]], dispatch: DispatchMeta, overlap_degree: int = 2,
                   ) -> tuple[list[CommMeta], list[CalcMeta]]:
    """Plan for ALL ranks: returns (comm_metas[rank], calc_metas[rank])."""
    cs, cp = dispatch.chunk_size, dispatch.cp_size
    n_chunks = dispatch.num_chunks
    owner = np.empty(n_chunks, dtype=np.int64)
    local_off = np.empty(n_chunks, dtype=np.int64)
    for r, part in enumerate(dispatch.partitions):
        for i, ch in enumerate(part):
            owner[ch], local_off[ch] = r, i * cs
    overlap_degree = max(1, overlap_degree)

    def k_runs(c: int, e: int):
        """Cut global cols [c,e) into maximal runs with one owner and contiguous local layout:
        yields (g0, g1, owner, local_start)."""
        ch = c // cs
        while c < e:
            o = int(owner[ch])
            g0, l0 = c, int(local_off[ch]) + (c - ch * cs)
            g1 = min(e, (ch + 1) * cs)
            nxt = ch + 1
            while g1 < e and nxt < n_chunks and int(owner[nxt]) == o and int(local_off[nxt]) == int(local_off[nxt - 1]) + cs:
                g1 = min(e, (nxt + 1) * cs)
                nxt += 1
            yield g0, g1, o, l0
            c, ch = g1, nxt

    # ---- pass 1: per rank, the pieces (slice x host span x owner run) and the remote requests
    pieces_per_rank = []   # rank -> list of (slice, a, b, q_local_base, g0, g1, owner, k_local_start)
    requests = []          # rank -> owner -> list of (g0, g1, area)
    for r in range(cp):
        spans = dispatch.host_spans(r)
        pieces, req = [], {}
        for s in slices:
            q0, q1, k0, k1, t = s
            if q1 <= q0 or k1 <= k0:
                continue
            for (h0, h1, hl) in spans:
                a, b = max(q0, h0), min(q1, h1)
                if a >= b:
                    continue
                c_min, c_max = max(_lo(s, a), k0), min(_hi(s, b - 1), k1)
                if c_max <= c_min:
                    continue
                for g0, g1, o, l0 in k_runs(c_min, c_max):
                    pieces.append((s, a, b, hl - h0, g0, g1, o, l0))
                    if o != r:
                        req.setdefault(o, []).append((g0, g1, slice_area(s, a, b, g0, g1)))
        pieces_per_rank.append(pieces)
        requests.append(req)

    # ---- pass 2: per rank, merge requests into segments, split into stages by area
    seg_tables = []  # rank -> list per stage of [(g0, g1, owner, buf_off)]
    for r in range(cp):
        segs = []
        for o, lst in requests[r].items():
            lst = sorted(lst)
            cur0, cur1, area = lst[0]
            for g0, g1, ar in lst[1:]:
                # merge only what is contiguous on the owner (same owner, touching or overlapping global
                # spans inside one local run)
                if g0 <= cur1 and _same_run(owner, local_off, cs, cur0, g1):
                    cur1, area = max(cur1, g1), area + ar
                else:
                    segs.append((cur0, cur1, o, area))
                    cur0, cur1, area = g0, g1, ar
            segs.append((cur0, cur1, o, area))
        segs.sort()
        total_area = sum(x[3] for x in segs)
        stages: list[list] = [[] for _ in range(overlap_degree)]
        acc, st = 0, 0
        for g0, g1, o, ar in segs:
            # advance to the stage whose area quota this segment starts in
            while st < overlap_degree - 1 and acc >= (st + 1) * total_area / overlap_degree:
                st += 1
            stages[st].append((g0, g1, o))
            acc += ar
        table = []
        for lst in stages:
            off, rows = 0, []
            for g0, g1, o in lst:
                rows.append((g0, g1, o, off))
                off += g1 - g0
            table.append(rows)
        seg_tables.append(table)

    # ---- pass 3: kernel slices per rank / stage
    calc_metas = []
    for r in range(cp):
        local = AttnArg()
        remotes = [AttnArg() for _ in range(overlap_degree)]
        seg_index = {}  # (stage) -> sorted list for lookup
        for st, rows in enumerate(seg_tables[r]):
            for g0, g1, o, off in rows:
                seg_index.setdefault(o, []).append((g0, g1, st, off))
        for (s, a, b, qshift, g0, g1, o, l0) in pieces_per_rank[r]:
            if o == r:
                target, kshift = local, l0 - g0
                _emit(target, s, a, b, g0, g1, qshift, kshift)
            else:
                for (sg0, sg1, st, off) in seg_index[o]:
                    x0, x1 = max(g0, sg0), min(g1, sg1)
                    if x0 < x1:
                        _emit(remotes[st], s, a, b, x0, x1, qshift, off - sg0)
        calc_metas.append(CalcMeta(local, remotes, dispatch.shard_seqlen, dispatch.shard_seqlen,
                                   [sum(g1 - g0 for g0, g1, _, _ in rows) for rows in seg_tables[r]]))

    # ---- pass 4: transfer tables per stage
    comm_metas = [CommMeta([], []) for _ in range(cp)]
    shard = dispatch.shard_seqlen
    for st in range(overlap_degree):
        # sender side: elementary splits of every owner's local rows
        cuts = [set([0, shard]) for _ in range(cp)]
        wants: list[list[tuple[int, int, int]]] = [[] for _ in range(cp)]  # owner -> (l0, l1, dst)
        for d in range(cp):
            for g0, g1, o, _ in seg_tables[d][st]:
                l0 = _to_local(local_off, cs, g0)
                l1 = l0 + (g1 - g0)
                cuts[o].update((l0, l1))
                wants[o].append((l0, l1, d))
        splits = []  # owner -> list of (l0, l1, [dsts])
        for o in range(cp):
            pts = sorted(cuts[o])
            sp = []
            for l0, l1 in zip(pts[:-1], pts[1:]):
                dsts = sorted({d for (w0, w1, d) in wants[o] if w0 <= l0 and l1 <= w1})
                sp.append((l0, l1, dsts))
            splits.append(sp)
        for r in range(cp):
            in_sizes = [l1 - l0 for l0, l1, _ in splits[r]]
            dst_idx = [d for _, _, d in splits[r]]
            out_sizes, src_idx = [], []
            for g0, g1, o, _ in seg_tables[r][st]:
                l0 = _to_local(local_off, cs, g0)
                l1 = l0 + (g1 - g0)
                for s0, s1, dsts in splits[o]:
                    if l0 <= s0 and s1 <= l1:
                        assert r in dsts
                        out_sizes.append(s1 - s0)
                        src_idx.append(o)
            comm_metas[r].kv_group_collective_args_list.append(GroupCollectiveArg(in_sizes, out_sizes, dst_idx, src_idx))
            comm_metas[r].num_remote_kv_tokens_per_stage.append(sum(out_sizes))
    return comm_metas, calc_metas


def _to_local(local_off, cs, g):
    ch = g // cs
    return int(local_off[ch]) + (g - ch * cs)


def _same_run(owner, local_off, cs, g0, g1) -> bool:
    """True if global cols [g0,g1) lie in one owner run with contiguous local layout."""
    c0, c1 = g0 // cs, (g1 - 1) // cs
    for ch in range(c0 + 1, c1 + 1):
        if owner[ch] != owner[c0] or local_off[ch] != local_off[ch - 1] + cs:
            return False
    return True


def _emit(arg: AttnArg, s, a, b, c, e, qshift, kshift) -> None:
    for (x, y, k0, k1, t) in _restrict(s, a, b, c, e):
        arg.q_ranges.append((x + qshift, y + qshift))
        arg.k_ranges.append((k0 + kshift, k1 + kshift))
        arg.attn_type_map.append(t)
        arg.total_area += slice_area((x, y, k0, k1, t))
