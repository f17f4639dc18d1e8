"""GPU parity at the FULL sizes of the named BASELINE configs (bench_workloads.py), through the C ABI, against the oracle on
sampled sub-problems that the domain makes exact:

* C2 (64K packed varlen, causal + sliding-window documents): documents are independent, so out / lse / dq / dk / dv of a
  sampled document equal the oracle run on that document alone.
* C5 (1M, 64:8 GQA, block-sparse FULL rectangles, cp=8): ONE rank's share, planned by the real planner (dispatch + stages),
  executed stage by stage with the real kernels (accumulate forward, global-lse backward); the remote K/V of every stage come
  from the oracle's naive group_cast.  Sampled q blocks: out / lse / dq vs dense attention over exactly the key blocks the
  mask gives them; this rank's partial dk / dv of a sampled key block vs the oracle over the rank's q rows that see it.
"""
import pytest
import torch

from bench_workloads import make_workload
from oracle import ffa_oracle as orc

pytestmark = pytest.mark.gpu
D = 128


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def test_c2_full_size_sampled_documents():
    _need_gpu()
    from magiattention_b200 import _lib

    dev = torch.device("cuda:0")
    w = make_workload("C2")
    T, H = w["seq_len"], 8  # 8 of the 32 heads: the kernels treat heads independently
    torch.manual_seed(5)
    q, k, v, do = (torch.randn(T, H, D, dtype=torch.bfloat16, device=dev) for _ in range(4))
    mk = lambda r, shape: torch.tensor(r, dtype=torch.int32, device=dev).reshape(shape)
    qr, kr, ty = mk(w["q_ranges"], (-1, 2)), mk(w["k_ranges"], (-1, 2)), mk(w["types"], (-1,))
    out = torch.empty_like(q)
    lse = torch.empty(T, H, dtype=torch.float32, device=dev)
    scale = D ** -0.5
    _lib.ffa_fwd(q, k, v, out, lse, qr, kr, ty, scale)
    dq = torch.zeros(T, H, D, dtype=torch.float32, device=dev)
    dk, dv = torch.empty_like(k), torch.empty_like(v)
    _lib.ffa_bwd(do, q, k, v, out, lse, dq, dk, dv, qr, kr, ty, scale)
    # documents = maximal groups of slices sharing a k start (a causal doc is 1 slice, a sliding-window doc 2)
    docs = {}
    for (q0, q1), (k0, k1), t in zip(w["q_ranges"], w["k_ranges"], w["types"]):
        a = min(q0, k0)
        key = next((d for d in docs if d[0] <= a < d[1]), None)
        if key is None:
            docs[(a, max(q1, k1))] = [((q0, q1), (k0, k1), t)]
        else:
            lst = docs.pop(key)
            docs[(key[0], max(key[1], q1, k1))] = lst + [((q0, q1), (k0, k1), t)]
    assert sum(b - a for a, b in docs) == T
    # a few documents of each kind, bounded so the dense oracle stays cheap
    picked, kinds = [], {1: 0, 2: 0}
    for (a, b), sl in sorted(docs.items(), key=lambda x: -(x[0][1] - x[0][0])):
        if b - a <= 4096 and kinds[len(sl)] < 3 and b - a >= 300:
            picked.append(((a, b), sl)); kinds[len(sl)] += 1
    assert kinds[1] >= 1 and kinds[2] >= 1, "need causal and sliding-window documents in the sample"
    for (a, b), sl in picked:
        qrs = [[x[0][0] - a, x[0][1] - a] for x in sl]
        krs = [[x[1][0] - a, x[1][1] - a] for x in sl]
        tys = [x[2] for x in sl]
        mask = orc.mask_from_slices(qrs, krs, tys, b - a, b - a).to(dev)
        qd, kd, vd, dod = q[a:b], k[a:b], v[a:b], do[a:b]
        ro, rl = orc.attn_fwd(qd, kd, vd, mask, scale, high_precision=True)
        lo, ll = orc.attn_fwd(qd, kd, vd, mask, scale, high_precision=False)
        orc.check_close(f"doc[{a},{b}) out", out[a:b], ro, lo, "out")
        orc.check_close(f"doc[{a},{b}) lse", lse[a:b], rl, ll, "lse")
        ref = orc.attn_bwd(dod, qd, kd, vd, ro.double(), rl.double(), mask, scale, high_precision=True)
        lp = orc.attn_bwd(dod, qd, kd, vd, lo, ll, mask, scale, high_precision=False)
        for nm, g, r_, l_ in zip(("dq", "dk", "dv"), (dq[a:b], dk[a:b], dv[a:b]), ref, lp):
            orc.check_close(f"doc[{a},{b}) {nm}", g, r_, l_, nm, atol=orc.grad_atol(r_))


def test_c5_one_rank_share_real_kernels():
    _need_gpu()
    from magiattention_b200 import _lib
    from magiattention_b200.functional.dist_attn import DistAttnRuntime
    from magiattention_b200.meta.planner import make_attn_meta, make_dispatch_meta

    dev = torch.device("cuda:0")
    w = make_workload("C5")
    S, HQ, HK, cp, degree, rank = w["seq_len"], 16, 2, 8, 4, 3  # 16:2 heads keep the 8:1 GQA ratio at a quarter of the memory
    sl = [(q_[0], q_[1], k_[0], k_[1], t) for q_, k_, t in zip(w["q_ranges"], w["k_ranges"], w["types"])]
    dm = make_dispatch_meta(sl, S, 2048, cp)
    comm, calc = make_attn_meta(sl, dm, degree)
    n_slices = len(calc[rank].local_attn_arg.q_ranges) + sum(len(a.q_ranges) for a in calc[rank].remote_attn_args_list)
    assert n_slices > 400, n_slices
    torch.manual_seed(9)
    kg, vg = (torch.randn(S, HK, D, dtype=torch.bfloat16, device=dev) for _ in range(2))
    pos = [dm.position_ids(r).to(dev) for r in range(cp)]
    ql, dol = (torch.randn(dm.shard_seqlen, HQ, D, dtype=torch.bfloat16, device=dev) for _ in range(2))
    kl, vl = [kg[p] for p in pos], [vg[p] for p in pos]
    scale = D ** -0.5
    rt = DistAttnRuntime(comm[rank], calc[rank], None)
    out_acc = torch.zeros(ql.shape, dtype=torch.float32, device=dev)
    lse_acc = torch.full((ql.shape[0], HQ), float("-inf"), dtype=torch.float32, device=dev)
    stage_kv = []
    for st in range(degree):
        args = [comm[r].kv_group_collective_args_list[st] for r in range(cp)]
        kw = dict(input_split_sizes=[a.input_split_size_list for a in args], dst_indices=[a.dst_indices_list for a in args],
                  src_index=[a.src_index_list for a in args], output_split_sizes=[a.output_split_size_list for a in args])
        stage_kv.append((orc.group_cast_ref(kl, **kw)[rank].contiguous(), orc.group_cast_ref(vl, **kw)[rank].contiguous()))
    rt.attn_fwd(ql, kl[rank].contiguous(), vl[rank].contiguous(), calc[rank].local_attn_arg, out_acc, lse_acc, scale, 0.0)
    for st in range(degree):
        rt.attn_fwd(ql, stage_kv[st][0], stage_kv[st][1], calc[rank].remote_attn_args_list[st], out_acc, lse_acc, scale, 0.0)
    out_l = out_acc.to(torch.bfloat16)
    dq = torch.zeros(ql.shape, dtype=torch.float32, device=dev)
    dk_loc = torch.zeros(dm.shard_seqlen, HK, D, dtype=torch.float32, device=dev)
    dv_loc = torch.zeros_like(dk_loc)
    rt.attn_bwd(dol, ql, kl[rank].contiguous(), vl[rank].contiguous(), out_l, lse_acc, calc[rank].local_attn_arg, dq, dk_loc, dv_loc,
                scale, 0.0)
    parts = []
    for st in range(degree):
        rk, rv = stage_kv[st]
        a = calc[rank].remote_attn_args_list[st].to_ffa_args(dev)
        dkp, dvp = torch.zeros(rk.shape, dtype=torch.float32, device=dev), torch.zeros(rv.shape, dtype=torch.float32, device=dev)
        _lib.ffa_bwd(dol, ql, rk, rv, out_l, lse_acc, dq, dkp, dvp, a["q_ranges"], a["k_ranges"], a["attn_type_map"], scale)
        parts.append((dkp, dvp))
    torch.cuda.synchronize()

    # ---- oracle on sampled q blocks of this rank: keys = exactly the rectangles of the block
    blocks = sorted({tuple(q_) for q_ in w["q_ranges"]})
    keys_of = {}
    for q_, k_ in zip(w["q_ranges"], w["k_ranges"]):
        keys_of.setdefault(tuple(q_), []).append(tuple(k_))
    inv = torch.full((S,), -1, dtype=torch.int64, device=dev)
    inv[pos[rank]] = torch.arange(dm.shard_seqlen, device=dev)
    def local_contig(b):
        r = inv[b[0]:b[1]]
        return bool((r >= 0).all()) and bool((r[1:] - r[:-1] == 1).all())

    mine = [b for b in blocks if b[1] - b[0] <= 2048 and local_contig(b)]
    assert len(mine) >= 3
    g = HQ // HK
    for (a, b) in (mine[1], mine[len(mine) // 2], mine[-1]):
        cols = torch.cat([torch.arange(c0, c1, device=dev) for c0, c1 in sorted(set(keys_of[(a, b)]))])
        rows = inv[a:b]
        qd, dod = ql[rows], dol[rows]
        mask = torch.ones(b - a, cols.numel(), dtype=torch.bool, device=dev)
        ro, rl = orc.attn_fwd(qd, kg[cols], vg[cols], mask, scale, high_precision=True)
        lo, ll = orc.attn_fwd(qd, kg[cols], vg[cols], mask, scale, high_precision=False)
        orc.check_close(f"q block [{a},{b}) out", out_l[rows], ro, lo, "out")
        orc.check_close(f"q block [{a},{b}) lse", lse_acc[rows], rl, ll, "lse")
        rdq = orc.attn_bwd(dod, qd, kg[cols], vg[cols], ro.double(), rl.double(), mask, scale, high_precision=True)[0]
        ldq = orc.attn_bwd(dod, qd, kg[cols], vg[cols], lo, ll, mask, scale, high_precision=False)[0]
        orc.check_close(f"q block [{a},{b}) dq", dq[rows], rdq, ldq, "dq", atol=orc.grad_atol(rdq))
    # ---- this rank's LOCAL-stage dk / dv of a sampled (non-global) key block that it owns: every local q row that sees the
    # block contributes there; oracle = partial backward with the global (out, lse) of those rows over the block's keys
    bi = {b: i for i, b in enumerate(blocks)}
    cand = [b for b in mine if bi[b] >= 3]
    (a, b) = cand[len(cand) // 2]
    j = bi[(a, b)]
    q_rows = torch.cat([inv[blocks[i][0]:blocks[i][1]] for i in range(j, min(j + 3, len(blocks)))])
    q_rows = q_rows[q_rows >= 0]
    krows = inv[a:b]
    mask = torch.ones(q_rows.numel(), b - a, dtype=torch.bool, device=dev)
    args = (dol[q_rows], ql[q_rows], kg[a:b], vg[a:b], out_l[q_rows].double(), lse_acc[q_rows].double(), mask, scale)
    _, rdk, rdv = orc.attn_bwd(*args, high_precision=True)
    _, ldk, ldv = orc.attn_bwd(args[0], args[1], args[2], args[3], out_l[q_rows], lse_acc[q_rows], mask, scale, high_precision=False)
    orc.check_close(f"k block [{a},{b}) local dk", dk_loc[krows], rdk, ldk, "dk", atol=orc.grad_atol(rdk))
    orc.check_close(f"k block [{a},{b}) local dv", dv_loc[krows], rdv, ldv, "dv", atol=orc.grad_atol(rdv))
    assert g == 8
