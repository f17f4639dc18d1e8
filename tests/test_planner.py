"""CPU tests of the host-side CP planner: for several masks and cp sizes, simulate every rank's staged
attention with the oracle on the planner's local slices + transfer tables and compare with the global
oracle result (forward and the dK/dV reduction path)."""
import pytest
import torch

from magiattention_b200.meta.planner import _restrict, make_attn_meta, make_dispatch_meta, slice_area
from mask_configs import mask_configs
from oracle import ffa_oracle as orc


def _slices(cfg):
    return [(q[0], q[1], k[0], k[1], t) for q, k, t in zip(cfg["q_ranges"], cfg["k_ranges"], cfg["types"])]


@pytest.mark.parametrize("t", [0, 1, 2, 3])
def test_restrict_exact(t):
    """Band decomposition reproduces the restricted mask exactly, for random sub-rectangles."""
    g = torch.Generator().manual_seed(t)
    for _ in range(60):
        q0 = int(torch.randint(0, 20, (1,), generator=g)); sq = int(torch.randint(1, 40, (1,), generator=g))
        k0 = int(torch.randint(0, 20, (1,), generator=g)); sk = int(torch.randint(1, 40, (1,), generator=g))
        s = (q0, q0 + sq, k0, k0 + sk, t)
        a = int(torch.randint(0, 30, (1,), generator=g)); b = a + int(torch.randint(1, 40, (1,), generator=g))
        c = int(torch.randint(0, 30, (1,), generator=g)); e = c + int(torch.randint(1, 40, (1,), generator=g))
        N = 100
        full = orc.mask_from_slices([[s[0], s[1]]], [[s[2], s[3]]], [t], N, N)
        want = torch.zeros_like(full)
        want[a:b, c:e] = full[a:b, c:e]
        bands = _restrict(s, a, b, c, e)
        got = torch.zeros_like(full)
        for (x, y, kk0, kk1, tt) in bands:
            m = orc.mask_from_slices([[x, y]], [[kk0, kk1]], [tt], N, N)
            assert not (got & m).any(), "bands overlap"
            got |= m
        assert torch.equal(got, want), (s, a, b, c, e, bands)
        assert sum(slice_area(bd) for bd in bands) == int(want.sum()) == slice_area(s, a, b, c, e)


def _simulate(cfg_slices, T, cp, chunk, degree, H=2, HK=1, D=16, seed=0):
    torch.manual_seed(seed)
    q = torch.randn(T, H, D, dtype=torch.float64)
    k = torch.randn(T, HK, D, dtype=torch.float64)
    v = torch.randn(T, HK, D, dtype=torch.float64)
    do = torch.randn(T, H, D, dtype=torch.float64)
    mask = orc.mask_from_slices([[s[0], s[1]] for s in cfg_slices], [[s[2], s[3]] for s in cfg_slices],
                                [s[4] for s in cfg_slices], T, T)
    ref_o, ref_l = orc.attn_fwd(q, k, v, mask)
    ref_dq, ref_dk, ref_dv = orc.attn_bwd(do, q, k, v, ref_o, ref_l, mask)
    dm = make_dispatch_meta(cfg_slices, T, chunk, cp)
    assert sorted(c for p in dm.partitions for c in p) == list(range(T // chunk))
    comm, calc = make_attn_meta(cfg_slices, dm, degree)
    pos = [dm.position_ids(r) for r in range(cp)]
    ql, kl, vl, dol = ([x[p] for p in pos] for x in (q, k, v, do))
    total_area = sum(cm.local_attn_arg.total_area + sum(a.total_area for a in cm.remote_attn_args_list) for cm in calc)
    assert total_area == int(mask.sum()), "planner slices must tile the mask exactly"
    dk_acc = [torch.zeros_like(x) for x in kl]
    dv_acc = [torch.zeros_like(x) for x in vl]
    outs, lses, dqs = [], [], []
    # stage buffers through the naive world-level group_cast
    stage_k, stage_v = [], []
    for st in range(degree):
        args = [comm[r].kv_group_collective_args_list[st] for r in range(cp)]
        kw = dict(input_split_sizes=[a.input_split_size_list for a in args],
                  dst_indices=[a.dst_indices_list for a in args], src_index=[a.src_index_list for a in args],
                  output_split_sizes=[a.output_split_size_list for a in args])
        for a in args:
            assert sum(a.input_split_size_list) == dm.shard_seqlen
        stage_k.append(orc.group_cast_ref(kl, **kw))
        stage_v.append(orc.group_cast_ref(vl, **kw))
    for r in range(cp):
        cm = calc[r]
        S = dm.shard_seqlen

        def part(arg, kk, vv):
            m = orc.mask_from_slices(arg.q_ranges, arg.k_ranges, arg.attn_type_map, S, kk.shape[0])
            return m, orc.attn_fwd(ql[r], kk, vv, m)

        m0, (o, l) = part(cm.local_attn_arg, kl[r], vl[r])
        masks = [m0]
        for st in range(degree):
            m, (o2, l2) = part(cm.remote_attn_args_list[st], stage_k[st][r], stage_v[st][r])
            masks.append(m)
            o, l = orc.merge_out_lse(o, l, o2, l2)
        torch.testing.assert_close(o, ref_o[pos[r]], atol=1e-10, rtol=1e-10)
        fin = ~torch.isinf(ref_l[pos[r]])
        assert torch.equal(torch.isinf(l), ~fin)
        torch.testing.assert_close(l[fin], ref_l[pos[r]][fin], atol=1e-10, rtol=1e-10)
        outs.append(o); lses.append(l)
        # backward with global (o, l): local part + per-stage partials
        dq, dk0, dv0 = orc.attn_bwd(dol[r], ql[r], kl[r], vl[r], o, l, masks[0])
        dk_acc[r] += dk0; dv_acc[r] += dv0
        parts_k, parts_v = [], []
        for st in range(degree):
            dq2, dkp, dvp = orc.attn_bwd(dol[r], ql[r], stage_k[st][r], stage_v[st][r], o, l, masks[1 + st])
            dq = dq + dq2
            parts_k.append(dkp); parts_v.append(dvp)
        dqs.append((dq, parts_k, parts_v))
        torch.testing.assert_close(dq, ref_dq[pos[r]], atol=1e-9, rtol=1e-9)
    for st in range(degree):
        args = [comm[r].kv_group_collective_args_list[st].to_group_reduce_args() for r in range(cp)]
        kw = dict(input_split_sizes=[a["input_split_sizes"] for a in args], dst_index=[a["dst_index"] for a in args],
                  src_indices=[a["src_indices"] for a in args], output_split_sizes=[a["output_split_sizes"] for a in args])
        dk_acc = orc.group_reduce_ref([dqs[r][1][st] for r in range(cp)], dk_acc, **kw)
        dv_acc = orc.group_reduce_ref([dqs[r][2][st] for r in range(cp)], dv_acc, **kw)
    for r in range(cp):
        torch.testing.assert_close(dk_acc[r], ref_dk[pos[r]], atol=1e-9, rtol=1e-9)
        torch.testing.assert_close(dv_acc[r], ref_dv[pos[r]], atol=1e-9, rtol=1e-9)
    return dm, comm, calc


@pytest.mark.parametrize("name", ["causal", "full", "varlen_causal_1k", "sliding_window_causal", "all_types_aspects",
                                  "sparse_attn_2k", "uncovered_rows", "varlen_block_causal_2k"])
@pytest.mark.parametrize("cp,degree", [(2, 1), (4, 2), (8, 3)])
def test_planner_matches_global_oracle(name, cp, degree):
    cfg = next(c for c in mask_configs(4) if c["name"] == name)
    T = cfg["seqlen"]
    chunk = 16
    pad = (-T) % (chunk * cp)
    _simulate(_slices(cfg), T + pad, cp, chunk, degree)


def test_causal_dispatch_is_balanced_and_zero_redundant():
    """C4-like shape scaled down: causal, cp=8: equal area per rank, and every rank fetches only keys it can see
    (reference planner finding, SURVEY App. A: per-rank area = 1.0000 x mean, remote rows < S)."""
    T, cp, chunk = 4096, 8, 32
    s = [(0, T, 0, T, 1)]
    dm = make_dispatch_meta(s, T, chunk, cp)
    comm, calc = make_attn_meta(s, dm, 2)
    areas = [cm.local_attn_arg.total_area + sum(a.total_area for a in cm.remote_attn_args_list) for cm in calc]
    assert max(areas) - min(areas) <= 0.002 * (sum(areas) / cp)
    assert sum(areas) == T * (T + 1) // 2
    for r in range(cp):
        rows = sum(comm[r].num_remote_kv_tokens_per_stage)
        assert rows <= T - dm.shard_seqlen
        # remote slices of a causal mask are FULL except where a remote chunk holds the diagonal (never: own chunk)
        for a in calc[r].remote_attn_args_list:
            assert all(t == 0 for t in a.attn_type_map)


def test_sliding_window_inference_known_answer():
    """SURVEY App. A.2 known answer: cu_seqlens [0,3000,3500,8192], window (1024,0) -> area 6 959 950."""
    from magiattention_b200.api import infer_attn_mask_from_cu_seqlens

    qr, kr, ty, tq, tk = infer_attn_mask_from_cu_seqlens([0, 3000, 3500, 8192], [0, 3000, 3500, 8192], False, (1024, 0))
    assert tq == tk == 8192
    ints = [t.to_int_type() for t in ty]
    assert orc.mask_area(qr.to_naive_ranges(), kr.to_naive_ranges(), ints) == 6959950
    # brute-force band mask
    m = orc.mask_from_slices(qr.to_naive_ranges(), kr.to_naive_ranges(), ints, 8192, 8192)
    i = torch.arange(8192).unsqueeze(1); j = torch.arange(8192).unsqueeze(0)
    doc = torch.bucketize(torch.arange(8192), torch.tensor([3000, 3500, 8192]), right=True)
    want = (doc.unsqueeze(1) == doc.unsqueeze(0)) & (j <= i) & (j >= i - 1024)
    assert torch.equal(m, want)


def test_planner_block_sparse_many_slices():
    """C5-like (scaled down): varlen block-causal rectangles (own block + previous two + global first block), cp=8."""
    import random

    random.seed(3)
    T, cp, chunk = 2048, 8, 16
    edges = [0]
    while edges[-1] < T:
        edges.append(min(T, edges[-1] + random.choice([16, 32, 48, 64, 96])))
    sl = []
    for i in range(len(edges) - 1):
        lo = max(i - 2, 0)
        sl.append((edges[i], edges[i + 1], edges[lo], edges[i + 1], 0))
        if lo > 0:
            sl.append((edges[i], edges[i + 1], edges[0], edges[1], 0))
    dm, comm, calc = _simulate(sl, T, cp, chunk, 2, H=2, HK=1, D=8)
    # zero-redundancy: nobody fetches more remote rows than exist outside its shard
    for r in range(cp):
        assert sum(comm[r].num_remote_kv_tokens_per_stage) <= T - dm.shard_seqlen


def test_sliding_window_inference_bruteforce():
    """infer_attn_mask_from_sliding_window vs a brute-force band mask (flash-attn convention, bottom-right aligned),
    for random q/k spans and windows incl. unbounded sides (-1) — the cases of the reference's
    tests/test_api/test_functools.py (known answers there are slice lists; here the dense masks must agree)."""
    from magiattention_b200.api import infer_attn_mask_from_sliding_window
    from magiattention_b200.common import AttnRange

    g = torch.Generator().manual_seed(5)
    ri = lambda lo, hi: int(torch.randint(lo, hi, (1,), generator=g))
    for _ in range(200):
        q0, sq, k0, sk = ri(0, 10), ri(1, 40), ri(0, 10), ri(1, 40)
        wl = ri(-1, 45)
        wr = ri(-1, 45)
        qr, kr, ty = infer_attn_mask_from_sliding_window(AttnRange(q0, q0 + sq), AttnRange(k0, k0 + sk), (wl, wr))
        N = 64
        got = orc.mask_from_slices(qr.to_naive_ranges(), kr.to_naive_ranges(), [t.to_int_type() for t in ty], N, N)
        i = torch.arange(sq).unsqueeze(1); j = torch.arange(sk).unsqueeze(0)
        d = sk - sq
        band = torch.ones(sq, sk, dtype=torch.bool)
        if wl >= 0:
            band &= j >= i + d - wl
        if wr >= 0:
            band &= j <= i + d + wr
        if sq > sk:
            band &= i >= sq - sk  # reference semantics: rows above the bottom-right aligned square are dropped
        want = torch.zeros(N, N, dtype=torch.bool)
        want[q0:q0 + sq, k0:k0 + sk] = band
        assert torch.equal(got, want), (q0, sq, k0, sk, wl, wr)
        # slices are area-disjoint
        cover = torch.zeros(N, N, dtype=torch.int32)
        for (a, b), (c, e) in zip(qr.to_naive_ranges(), kr.to_naive_ranges()):
            cover[a:b, c:e] += 1
        assert int(cover.max()) <= 1


@pytest.mark.parametrize("alg", ["min_heap", "sequential", "sorted_sequential"])
@pytest.mark.parametrize("cp", [2, 4])
def test_dispatch_algorithms_partition_and_plan(alg, cp):
    """Every dispatch algorithm yields a partition of the chunks with equal counts, and the CP plan built on it reproduces
    the global mask (areas add up; simulated CP attention equals the oracle)."""
    from magiattention_b200.meta.planner import make_attn_meta, make_dispatch_meta, slice_area

    cfg = next(c for c in mask_configs(4) if c["name"] == "all_types_aspects")
    sl = [(q[0], q[1], k[0], k[1], t) for q, k, t in zip(cfg["q_ranges"], cfg["k_ranges"], cfg["types"])]
    chunk = 16
    T = cfg["seqlen"] + (-cfg["seqlen"]) % (chunk * cp)
    dm = make_dispatch_meta(sl, T, chunk, cp, alg)
    flat = sorted(c for p in dm.partitions for c in p)
    assert flat == list(range(T // chunk)) and len({len(p) for p in dm.partitions}) == 1
    if alg == "sequential":
        assert dm.partitions[0] == list(range(T // chunk // cp))
    comm, calc = make_attn_meta(sl, dm, 2)
    total = sum(c.local_attn_arg.total_area + sum(a.total_area for a in c.remote_attn_args_list) for c in calc)
    assert total == sum(slice_area(s) for s in sl)
    loads = [c.local_attn_arg.total_area + sum(a.total_area for a in c.remote_attn_args_list) for c in calc]
    if alg != "sequential":  # the balancing algorithms keep the heaviest rank near the mean
        assert max(loads) <= 1.35 * (sum(loads) / cp)


def test_config_classes_match_reference_surface():
    """Config objects as the reference's docs construct them (api/magi_attn_interface.py:1385-1395) are accepted and hashable;
    every knob either works or raises."""
    from magiattention_b200.common.enum import AttnOverlapMode
    from magiattention_b200.config import (BSDispatchAlg, DispatchConfig, DistAttnConfig, DPDispatchAlg, GreedyOverlapAlg,
                                           GrpCollConfig, LBDispatchAlg, MinHeapDispatchAlg, OverlapConfig, SequentialDispatchAlg,
                                           SortedSequentialSelectAlg, ToppHeapDispatchAlg, UniformOverlapAlg)

    c = DistAttnConfig(dispatch_config=DispatchConfig(chunk_size=512, alg=MinHeapDispatchAlg()),
                       overlap_config=OverlapConfig(mode=AttnOverlapMode.STATIC, degree=2, min_chunk_size=512,
                                                    max_num_chunks=64, alg=UniformOverlapAlg()),
                       grpcoll_config=GrpCollConfig())
    assert hash(c) == hash(DistAttnConfig(dispatch_config=DispatchConfig(chunk_size=512, alg=MinHeapDispatchAlg()),
                                          overlap_config=OverlapConfig(AttnOverlapMode.STATIC, 2)))
    assert c.dispatch_config.planner_alg == "min_heap" and c.overlap_config.num_stages(8) == 2
    assert DispatchConfig(alg=SequentialDispatchAlg()).planner_alg == "sequential"
    assert DispatchConfig(alg=SortedSequentialSelectAlg()).planner_alg == "sorted_sequential"
    for alg in (ToppHeapDispatchAlg(top_p=0.5), LBDispatchAlg(), DPDispatchAlg(), BSDispatchAlg()):
        with pytest.raises(NotImplementedError):
            DispatchConfig(alg=alg)
    with pytest.raises(NotImplementedError):
        OverlapConfig(degree=2, alg=GreedyOverlapAlg())
    # reference semantics of degree (overlap_solver.py:72-147): default 1, 0 = no overlap, None needs the dynamic mode
    assert OverlapConfig().degree == 1 and not OverlapConfig().enable and OverlapConfig(degree=2).enable
    z = OverlapConfig(degree=0)
    assert z.no_overlap and z.degree == 1 and z.max_num_chunks == 1 and z.num_stages(8) == 1
    dyn = OverlapConfig(mode=AttnOverlapMode.DYNAMIC, degree=None)
    assert dyn.num_stages(8) == 4 and dyn.num_stages(2) == 2 and OverlapConfig(degree=3).num_stages(1) == 1
    assert OverlapConfig(mode=AttnOverlapMode.DYNAMIC, degree=None, dynamic_max_degree=2).num_stages(8) == 2
    with pytest.raises(AssertionError):
        OverlapConfig(degree=None)  # static mode needs an explicit degree
    with pytest.raises(NotImplementedError):
        DispatchConfig(uneven_shard=True)


@pytest.mark.parametrize("seed", range(8))
def test_planner_random_slices_fuzz(seed):
    """Seeded fuzz of the whole planner: random area-disjoint slices of all four types (random grid cells), random cp,
    chunk size and stage count, through the full simulated CP forward/backward against the global oracle."""
    import numpy as np

    rng = np.random.default_rng(4321 + seed)
    cp = int(rng.choice([2, 3, 4, 8]))
    chunk = int(rng.choice([8, 16, 32]))
    T = chunk * cp * int(rng.integers(1, 5))
    degree = int(rng.integers(1, 4))

    def cuts(n):
        c = sorted(set([0, n] + [int(x) for x in rng.integers(1, max(n, 2), size=int(rng.integers(1, 6)))]))
        return [(a, b) for a, b in zip(c[:-1], c[1:]) if b > a]

    sl = []
    for (q0, q1) in cuts(T):
        for (k0, k1) in cuts(T):
            if rng.random() < 0.6:
                sl.append((q0, q1, k0, k1, int(rng.integers(0, 4))))
    _simulate(sl, T, cp, chunk, degree, seed=seed)


@pytest.mark.parametrize("cp,degree", [(2, 1), (4, 2), (8, 3)])
def test_pull_plan_equals_group_cast(cp, degree):
    """comm/nvl_exchange.make_pull_plan (the host side of the NVLink peer-memory exchange): pulling the listed row spans
    out of the owners' shards reproduces the receive buffer of the naive group_cast for every rank and stage."""
    from magiattention_b200.comm.nvl_exchange import make_pull_plan

    for cfg in mask_configs(4):
        sl = _slices(cfg)
        chunk = 16
        T = cfg["seqlen"] + (-cfg["seqlen"]) % (chunk * cp)
        dm = make_dispatch_meta(sl, T, chunk, cp)
        comm, _ = make_attn_meta(sl, dm, degree)
        kg = torch.arange(T * 2, dtype=torch.float64).reshape(T, 2)
        kl = [kg[dm.position_ids(r)] for r in range(cp)]
        for st in range(degree):
            args = [comm[r].kv_group_collective_args_list[st] for r in range(cp)]
            kw = dict(input_split_sizes=[a.input_split_size_list for a in args], dst_indices=[a.dst_indices_list for a in args],
                      src_index=[a.src_index_list for a in args], output_split_sizes=[a.output_split_size_list for a in args])
            want = orc.group_cast_ref(kl, **kw)
            for r in range(cp):
                out = torch.full_like(want[r], -1)
                for (s, s0, d0, n) in make_pull_plan(args, r):
                    assert s != r
                    out[d0:d0 + n] = kl[s][s0:s0 + n]
                assert torch.equal(out, want[r])
            # return path: every owner pulls the partial rows of its consumers and adds them == naive group_reduce
            from magiattention_b200.comm.nvl_exchange import make_reduce_pull_plan

            parts = [torch.randn(w.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(r + 7 * st))
                     for r, w in enumerate(want)]
            acc0 = [torch.zeros_like(x) for x in kl]
            rkw = [a.to_group_reduce_args() for a in args]
            ref = orc.group_reduce_ref(parts, [a.clone() for a in acc0], input_split_sizes=[a["input_split_sizes"] for a in rkw],
                                       dst_index=[a["dst_index"] for a in rkw], src_indices=[a["src_indices"] for a in rkw],
                                       output_split_sizes=[a["output_split_sizes"] for a in rkw])
            for o in range(cp):
                acc = acc0[o].clone()
                for (d, s0, l0, n) in make_reduce_pull_plan(args, o):
                    acc[l0:l0 + n] += parts[d][s0:s0 + n]
                torch.testing.assert_close(acc, ref[o], atol=1e-12, rtol=1e-12)


def test_dispatch_load_balance_matches_reference_planner():
    """Our min-heap dispatch against the partitions the reference's own planner (make_dispatch_meta_from_qk_ranges, MinHeap)
    emits for the same masks (fixture: tests/golden/dispatch_partitions.json, gen_dispatch_golden.py): equal chunk counts and
    a heaviest-rank mask area no worse than the reference's.  On the pure causal mask both give the zig-zag pairing."""
    import json
    import os

    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "dispatch_partitions.json")))
    assert len(cases) >= 18
    for c in cases:
        sl = [(q[0], q[1], k[0], k[1], t) for q, k, t in zip(c["q_ranges"], c["k_ranges"], c["types"])]
        T, cs, cp = c["total_seqlen"], c["chunk_size"], c["cp_size"]
        dm = make_dispatch_meta(sl, T, cs, cp)

        def rank_areas(parts):
            return [sum(slice_area(s, ch * cs, (ch + 1) * cs) for ch in p for s in sl if s[1] > s[0] and s[3] > s[2]) for p in parts]

        ours, ref = rank_areas(dm.partitions), rank_areas(c["partitions"])
        assert sum(ours) == sum(ref)
        assert sorted(len(p) for p in dm.partitions) == sorted(len(p) for p in c["partitions"])
        assert max(ours) <= max(ref) * 1.0001 + 1, (c["name"], cp, max(ours), max(ref))
        if c["name"] == "causal_64chunks":
            assert sorted(map(tuple, dm.partitions)) == sorted(map(tuple, c["partitions"]))


def test_c4_plan_matches_reference_planner_known_answers():
    """BASELINE config C4 (S = 262144 causal, cp = 8, chunk 2048): the numbers the reference's own planner produces for it
    (obtained by running DistAttnSolver on CPU): zig-zag partitions, zero-redundant remote KV volume of
    229376 (rank 0) ... 215040 (rank 7) rows, perfectly balanced mask area."""
    S, cp, cs = 262144, 8, 2048
    sl = [(0, S, 0, S, 1)]
    dm = make_dispatch_meta(sl, S, cs, cp)
    assert dm.partitions[0][:6] == [0, 15, 16, 31, 32, 47] and dm.partitions[7][:6] == [7, 8, 23, 24, 39, 40]
    comm, calc = make_attn_meta(sl, dm, 2)
    assert [sum(c.num_remote_kv_tokens_per_stage) for c in comm] == [229376 - 2048 * r for r in range(cp)]
    areas = [c.local_attn_arg.total_area + sum(a.total_area for a in c.remote_attn_args_list) for c in calc]
    assert len(set(areas)) == 1 and sum(areas) == S * (S + 1) // 2


# magi_attention/api/__init__.py __all__ of the reference (names only, snapshot taken by listing the module's __all__)
REFERENCE_API_ALL = ['magi_attn_varlen_key', 'magi_attn_varlen_dispatch', 'magi_attn_flex_key', 'magi_attn_flex_dispatch', 'dispatch', 'undispatch', 'roll', 'roll_simple', 'calc_attn', 'clear_cache', 'get_most_recent_key', 'get_position_ids', 'make_varlen_key_for_new_mask_after_dispatch', 'make_flex_key_for_new_mask_after_dispatch', 'flex_flash_attn_func', 'compute_pad_size', 'squash_batch_dim', 'infer_varlen_mask_from_batch', 'infer_attn_mask_from_sliding_window', 'infer_attn_mask_from_cu_seqlens', 'AttnForwardMeta', 'AttnMaskType', 'AttnOverlapMode', 'AttnRanges', 'DistAttnRuntimeKey', 'GeneralAttnMaskType', 'DistAttnConfig', 'DispatchConfig', 'OverlapConfig', 'GrpCollConfig', 'DispatchAlg', 'MinHeapDispatchAlg', 'ToppHeapDispatchAlg', 'SequentialDispatchAlg', 'SortedSequentialSelectAlg', 'LBDispatchAlg', 'DPDispatchAlg', 'BSDispatchAlg', 'OverlapAlg', 'UniformOverlapAlg', 'GreedyOverlapAlg', 'DistAttnRuntimeDictManager', 'dist_attn_runtime_dict_mgr']


def test_api_surface_covers_reference_all():
    """Every public name of the reference's ``magi_attention.api`` resolves in the alias package and in our own package."""
    import magi_attention.api as alias
    import magiattention_b200.api as ours

    missing = [n for n in REFERENCE_API_ALL if not hasattr(alias, n) or not hasattr(ours, n)]
    assert not missing, missing


@pytest.mark.parametrize("cfg,cp,degree", [("C3", 8, 4), ("C4", 8, 4), ("C5", 8, 4), ("C2", 4, 2)])
def test_named_workloads_plan_at_full_size(cfg, cp, degree):
    """The BASELINE configs at their FULL sizes (bench_workloads.py) through the planner, checked with size-independent
    properties: the ranks' slices (local + every stage) conserve the mask area exactly, every stage slice addresses rows inside
    its stage buffer, the stage buffers hold exactly the rows the transfer tables move, and the NVLink pull tables gather the
    same global token ids as the naive group_cast (int64 ids, so a wrong row anywhere is a mismatch)."""
    from bench_workloads import make_workload, slices_area
    from magiattention_b200.comm.nvl_exchange import make_pull_plan, make_reduce_pull_plan

    w = make_workload(cfg)
    S = w["seq_len"]
    sl = [(q_[0], q_[1], k_[0], k_[1], t) for q_, k_, t in zip(w["q_ranges"], w["k_ranges"], w["types"])]
    chunk = 2048 if S >= 65536 else 512
    assert S % (chunk * cp) == 0
    dm = make_dispatch_meta(sl, S, chunk, cp)
    comm, calc = make_attn_meta(sl, dm, degree)
    area = 0
    for r in range(cp):
        args = [calc[r].local_attn_arg] + list(calc[r].remote_attn_args_list)
        bounds = [dm.shard_seqlen] + [int(n) for n in comm[r].num_remote_kv_tokens_per_stage]
        for a, kmax in zip(args, bounds):
            qr, kr = list(a.q_ranges), list(a.k_ranges)
            area += slices_area(qr, kr, list(a.attn_type_map))
            assert slices_area(qr, kr, list(a.attn_type_map)) == a.total_area
            assert all(0 <= q0 <= q1 <= dm.shard_seqlen for q0, q1 in qr)
            assert all(0 <= k0 <= k1 <= kmax for k0, k1 in kr)
    assert area == w["area"], (area, w["area"])
    ids = [dm.position_ids(r).to(torch.int64) for r in range(cp)]
    assert torch.equal(torch.sort(torch.cat(ids)).values, torch.arange(S))
    for st in range(degree):
        args = [comm[r].kv_group_collective_args_list[st] for r in range(cp)]
        kw = dict(input_split_sizes=[a.input_split_size_list for a in args], dst_indices=[a.dst_indices_list for a in args],
                  src_index=[a.src_index_list for a in args], output_split_sizes=[a.output_split_size_list for a in args])
        want = orc.group_cast_ref([x[:, None] for x in ids], **kw)
        for r in range(cp):
            assert want[r].shape[0] == int(comm[r].num_remote_kv_tokens_per_stage[st])
            got = torch.full((want[r].shape[0],), -1, dtype=torch.int64)
            for (s, s0, d0, n) in make_pull_plan(args, r):
                assert s != r
                got[d0:d0 + n] = ids[s][s0:s0 + n]
            assert torch.equal(got, want[r][:, 0])
            # zero redundancy: a rank never fetches the same remote token twice in a stage
            assert torch.unique(got).numel() == got.numel()
            # the return path of the same stage lands every partial row on the owner's row of the same token
            for (d, s0, l0, n) in make_reduce_pull_plan(args, r):
                assert torch.equal(want[d][s0:s0 + n, 0], ids[r][l0:l0 + n])
