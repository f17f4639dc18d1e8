"""GPU parity tests of the flex-flash-attention kernels (through the C ABI) against the oracle.

Tolerance: the reference's two-sided rule (tests/test_attn/test_flex_flash_attn.py:680-738,839-873):
  ||x - ref64||_inf <= 2 * ||ref_lowprec - ref64||_inf   and   mismatch ratio (atol 1e-8; rtol out 0.05,
  lse 1e-3, dq 0.3, dk 0.15, dv 0.05) <= 2x the low-precision reference's own (capped at 0.75).
"""
import glob
import os

import numpy as np
import pytest
import torch

from mask_configs import mask_configs
from oracle import ffa_oracle as orc

pytestmark = pytest.mark.gpu
D = 128


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _run_case(qr, kr, ty, T, TK, HQ, HK, dtype, seed=0, check_bwd=True, D=D):
    from magiattention_b200.functional import flex_flash_attn_func

    dev = _dev()
    torch.manual_seed(seed)
    q = torch.randn(T, HQ, D, dtype=dtype, device=dev, requires_grad=True)
    k = torch.randn(TK, HK, D, dtype=dtype, device=dev, requires_grad=True)
    v = torch.randn(TK, HK, D, dtype=dtype, device=dev, requires_grad=True)
    do = torch.randn(T, HQ, D, dtype=dtype, device=dev)
    qr_t = torch.tensor(qr, dtype=torch.int32, device=dev).reshape(-1, 2)
    kr_t = torch.tensor(kr, dtype=torch.int32, device=dev).reshape(-1, 2)
    ty_t = torch.tensor(ty, dtype=torch.int32, device=dev)
    out, meta = flex_flash_attn_func(q, k, v, qr_t, kr_t, ty_t)
    mask = orc.mask_from_slices(qr, kr, ty, T, TK)
    qc, kc, vc, doc = (x.detach() for x in (q, k, v, do))  # oracle math runs on the GPU tensors (same code, fp64)
    ro, rl = orc.attn_fwd(qc, kc, vc, mask, high_precision=True)
    lo, ll = orc.attn_fwd(qc, kc, vc, mask, high_precision=False)
    ro64, _ = orc.attn_fwd(qc.double(), kc.double(), vc.double(), mask, high_precision=True)
    orc.check_close("out", out.detach(), ro64, lo, "out")
    orc.check_close("lse", meta.lse, rl, ll, "lse")
    if not check_bwd:
        return
    out.backward(do)
    ref = orc.attn_bwd(doc, qc.double(), kc.double(), vc.double(), ro64, rl.double(), mask, high_precision=True)
    lp = orc.attn_bwd(doc, qc, kc, vc, lo, ll, mask, high_precision=False)
    # gradients: scale-relative epsilon (see oracle.grad_atol): rows with a single visible key have exact-zero
    # true dq/dk, against which the reference's absolute 1e-8 epsilon would flag any fp32 rounding noise
    for name, g, r, l in zip(("dq", "dk", "dv"), (q.grad, k.grad, v.grad), ref, lp):
        orc.check_close(name, g, r, l, name, atol=orc.grad_atol(r))


@pytest.mark.parametrize("cfg", mask_configs(1), ids=[c["name"] for c in mask_configs(1)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
def test_ffa_mask_configs(cfg, dtype):
    """Every reference mask family (tests/mask_configs.py), GQA 2:1, fwd + bwd."""
    T = cfg["seqlen"]
    _run_case(cfg["q_ranges"], cfg["k_ranges"], cfg["types"], T, T, 2, 1, dtype)


@pytest.mark.parametrize("hq,hk", [(1, 1), (4, 4), (8, 2)])
def test_ffa_head_configs(hq, hk):
    _run_case([[0, 700]], [[0, 900]], [1], 700, 900, hq, hk, torch.bfloat16)


_HD64_CASES = ("varlen_causal_1k", "all_types_aspects", "sliding_window_causal", "uncovered_rows")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
def test_ffa_head_dim_64(dtype):
    """head_dim 64 (Llama-3-1B geometry of examples/torch_native): every mask type, GQA 4:1, ragged sizes, fwd + bwd."""
    cfgs = mask_configs(1)
    picked = [c for c in cfgs if c["name"] in _HD64_CASES]
    assert len(picked) == len(_HD64_CASES)
    for cfg in picked:
        T = cfg["seqlen"]
        _run_case(cfg["q_ranges"], cfg["k_ranges"], cfg["types"], T, T, 4, 1, dtype, D=64)
    _run_case([[0, 77], [77, 77], [100, 333], [333, 600]], [[0, 50], [3, 9], [10, 50], [20, 450]], [0, 0, 2, 3], 600, 450, 2, 2,
              dtype, D=64)
    _run_case([[0, 700]], [[0, 900]], [1], 700, 900, 8, 2, dtype, D=64)
    _run_case([], [], [], 130, 70, 2, 1, dtype, D=64)


def _run_sink_case(qr, kr, ty, T, TK, HQ, HK, dtype, layout, s_sink, D=D, seed=3):
    """fwd + bwd with an attention sink and return_max_logits against the oracle (same tolerance rule as _run_case)."""
    from magiattention_b200.functional import flex_flash_attn_func

    dev = _dev()
    torch.manual_seed(seed)
    q = torch.randn(T, HQ, D, dtype=dtype, device=dev, requires_grad=True)
    k = torch.randn(TK, HK, D, dtype=dtype, device=dev, requires_grad=True)
    v = torch.randn(TK, HK, D, dtype=dtype, device=dev, requires_grad=True)
    do = torch.randn(T, HQ, D, dtype=dtype, device=dev)
    sink = torch.randn((s_sink, HQ) if layout == "sh" else (T, s_sink, HQ), dtype=torch.float32, device=dev, requires_grad=True)
    mk = lambda x, shape: torch.tensor(x, dtype=torch.int32, device=dev).reshape(shape)
    out, meta = flex_flash_attn_func(q, k, v, mk(qr, (-1, 2)), mk(kr, (-1, 2)), mk(ty, (-1,)), sink=sink, sink_layout=layout,
                                     return_max_logits=True)
    mask = orc.mask_from_slices(qr, kr, ty, T, TK)
    qc, kc, vc, doc, sc = (x.detach() for x in (q, k, v, do, sink))
    ro64, rl = orc.attn_fwd(qc.double(), kc.double(), vc.double(), mask, sink=sc, sink_layout=layout)
    lo, ll = orc.attn_fwd(qc, kc, vc, mask, high_precision=False, sink=sc, sink_layout=layout)
    orc.check_close("out", out.detach(), ro64, lo, "out")
    orc.check_close("lse", meta.lse, rl.float(), ll, "lse")
    # max_logits: the kernel takes the max of fp32-accumulated 16-bit products, exactly what the fp64 oracle sees up to
    # fp32 accumulation error
    ml = orc.max_logits(qc.double(), kc.double(), mask)
    assert meta.max_logits.dtype == torch.float32 and meta.max_logits.shape == (HQ,)
    torch.testing.assert_close(meta.max_logits.double(), ml, atol=1e-4, rtol=1e-5)
    out.backward(do)
    ref = orc.attn_bwd(doc, qc.double(), kc.double(), vc.double(), ro64, rl.double(), mask)
    lp = orc.attn_bwd(doc, qc, kc, vc, lo, ll, mask, high_precision=False)
    for name, g, r, l in zip(("dq", "dk", "dv"), (q.grad, k.grad, v.grad), ref, lp):
        orc.check_close(name, g, r, l, name, atol=orc.grad_atol(r))
    ds_ref = orc.sink_bwd(sc, rl, ro64, doc, layout)
    ds_lp = orc.sink_bwd(sc, ll, lo, doc, layout)  # from the low-precision forward: the reference's error yardstick
    assert sink.grad is not None and sink.grad.dtype == torch.float32 and sink.grad.shape == sink.shape
    # the reference's dsink thresholds (test_flex_flash_attn.py:1760-1768): few elements => coarse mismatch floor.
    # "sh" dsink is a signed sum over all rows: an element that cancels to ~0 keeps the absolute error of its
    # summands (16-bit rounding of `out` inside delta), so atol is relative to the largest element.
    orc.check_close("dsink", sink.grad, ds_ref, ds_lp.float(), "dsink", atol=orc.dsink_atol(ds_ref),
                    min_mismatch=max(1.0 / max(ds_ref.numel(), 1), 8e-2), norm_floor=0.015)


@pytest.mark.parametrize("layout,s_sink", [("sh", 1), ("sh", 8), ("ssh", 3)])
def test_ffa_attention_sink_and_max_logits(layout, s_sink):
    """Attention sink (reference test_flex_flash_attn.py:1176-1200: fp32 sink, s_sink 1..8, layouts sh / ssh) + dsink +
    max_logits over mixed mask types, uncovered rows, ragged tails, GQA; d = 128 and d = 64."""
    cfgs = {c["name"]: c for c in mask_configs(1)}
    for name in ("all_types_aspects", "uncovered_rows"):
        c = cfgs[name]
        _run_sink_case(c["q_ranges"], c["k_ranges"], c["types"], c["seqlen"], c["seqlen"], 4, 2, torch.bfloat16, layout, s_sink)
    _run_sink_case([[0, 77], [100, 333]], [[0, 50], [10, 50]], [0, 2], 333, 50, 2, 2, torch.float16, layout, s_sink, D=64)
    _run_sink_case([], [], [], 130, 70, 2, 1, torch.bfloat16, layout, s_sink)  # sink only: out 0, lse = lse(sink), dsink 0


def test_sink_then_accumulate_stages():
    """CP stage order on one device: the host stage applies the sink, later stages merge their partial results into the
    same fp32 accumulators (dist_attn.py:1278-1297); backward per stage against the global (out, lse), dsink from the
    host stage only.  Must equal one-shot attention over the union of the key sets with the sink."""
    from magiattention_b200.functional.flex_flash_attn import _flex_flash_attn_backward, _flex_flash_attn_forward

    dev = _dev()
    torch.manual_seed(9)
    T, HQ, HK = 640, 4, 2
    q, do = (torch.randn(T, HQ, D, dtype=torch.bfloat16, device=dev) for _ in range(2))
    k, v = (torch.randn(T, HK, D, dtype=torch.bfloat16, device=dev) for _ in range(2))
    sink = torch.randn(5, HQ, dtype=torch.float32, device=dev)
    mk = lambda x, shape: torch.tensor(x, dtype=torch.int32, device=dev).reshape(shape)
    stages = [([[0, 640]], [[0, 256]], [0]), ([[100, 640]], [[256, 400]], [2]), ([[0, 500]], [[400, 640]], [1])]
    out = torch.zeros(T, HQ, D, dtype=torch.float32, device=dev)
    lse = torch.full((T, HQ), float("-inf"), dtype=torch.float32, device=dev)
    for i, (qr, kr, ty) in enumerate(stages):
        _flex_flash_attn_forward(q, k, v, sink if i == 0 else None, "sh", out, lse, mk(qr, (-1, 2)), mk(kr, (-1, 2)), mk(ty, (-1,)),
                                 D ** -0.5, 0.0, torch.float32)
    qr_all, kr_all, ty_all = (sum((s[i] for s in stages), []) for i in range(3))
    mask = orc.mask_from_slices(qr_all, kr_all, ty_all, T, T)
    ro, rl = orc.attn_fwd(q.double(), k.double(), v.double(), mask, sink=sink)
    lo, ll = orc.attn_fwd(q, k, v, mask, high_precision=False, sink=sink)
    orc.check_close("out", out, ro, lo, "out")
    orc.check_close("lse", lse, rl.float(), ll, "lse")
    out16 = out.to(q.dtype)
    dq = torch.zeros(T, HQ, D, dtype=torch.float32, device=dev)
    dk, dv = (torch.zeros(T, HK, D, dtype=torch.float32, device=dev) for _ in range(2))
    dsink = None
    for i, (qr, kr, ty) in enumerate(stages):
        _, _, _, ds = _flex_flash_attn_backward(do, q, k, v, sink if i == 0 else None, "sh", out16, lse, dq, dk, dv, None,
                                                mk(qr, (-1, 2)), mk(kr, (-1, 2)), mk(ty, (-1,)), D ** -0.5, 0.0)
        dsink = ds if i == 0 else dsink
        assert (ds is None) == (i != 0)
    ref = orc.attn_bwd(do, q.double(), k.double(), v.double(), ro, rl.double(), mask)
    lp = orc.attn_bwd(do, q, k, v, lo, ll, mask, high_precision=False)
    for name, g, r, l in zip(("dq", "dk", "dv"), (dq, dk, dv), ref, lp):
        orc.check_close(name, g, r, l, name, atol=orc.grad_atol(r))
    ds_ref, ds_lp = orc.sink_bwd(sink, rl, ro, do), orc.sink_bwd(sink, ll, lo, do)
    orc.check_close("dsink", dsink, ds_ref, ds_lp.float(), "dsink", atol=orc.dsink_atol(ds_ref),
                    min_mismatch=max(1.0 / ds_ref.numel(), 8e-2), norm_floor=0.015)


def test_correct_attn_out_lse_kernels():
    """functional.correct_attn_out_lse (b200_lse_merge) and correct_attn_out_lse_with_sink (b200_sink_merge) against the
    oracle's merge, out of place and in place, bf16 and fp32 partials."""
    from magiattention_b200.functional import correct_attn_out_lse, correct_attn_out_lse_with_sink

    dev = _dev()
    torch.manual_seed(2)
    T, H = 300, 4
    for dt, tol in ((torch.float32, 1e-5), (torch.bfloat16, 2e-2)):
        o1, o2 = (torch.randn(T, H, D, device=dev).to(dt) for _ in range(2))
        l1, l2 = (torch.randn(T, H, device=dev) for _ in range(2))
        l1[:7] = float("-inf"); o1[:7] = 0
        l2[4:11] = float("-inf"); o2[4:11] = 0
        ro, rl = orc.merge_out_lse(o1.double(), l1.double(), o2.double(), l2.double())
        o, l = correct_attn_out_lse(o1, l1, o2, l2)
        assert o.data_ptr() != o1.data_ptr() and torch.equal(torch.isinf(l), torch.isinf(rl))
        torch.testing.assert_close(o.double(), ro, atol=tol, rtol=tol)
        torch.testing.assert_close(l.double()[~torch.isinf(rl)], rl[~torch.isinf(rl)], atol=1e-5, rtol=1e-5)
        for layout, sink in (("sh", torch.randn(3, H, device=dev)), ("ssh", torch.randn(T, 3, H, device=dev))):
            ls = torch.logsumexp(orc._sink_hqs(sink.double(), T, layout), -1).t()
            so, sl = orc.merge_out_lse(ro, rl, torch.zeros_like(ro), ls)
            oc, lc = o.clone(), l.clone()
            o3, l3 = correct_attn_out_lse_with_sink(oc, lc, sink, layout, inplace=True)
            assert o3.data_ptr() == oc.data_ptr() and l3.data_ptr() == lc.data_ptr()
            torch.testing.assert_close(o3.double(), so, atol=2 * tol, rtol=2 * tol)
            torch.testing.assert_close(l3.double(), sl, atol=1e-5, rtol=1e-5)


def test_torch_custom_ops_equal_eager_path():
    """torch.ops.flex_flash_attn._flex_flash_attn_{forward,backward}_compilable (the reference's torch.compile seam,
    functional/custom_ops.py) give the same results as the eager launchers: fresh (0, -inf) / zero buffers in, merged results out."""
    import magiattention_b200.functional  # noqa: F401  (registers the ops)
    from magiattention_b200.functional.flex_flash_attn import _flex_flash_attn_backward, _flex_flash_attn_forward

    dev = _dev()
    torch.manual_seed(4)
    T, HQ, HK = 300, 2, 1
    q, do = (torch.randn(T, HQ, D, dtype=torch.bfloat16, device=dev) for _ in range(2))
    k, v = (torch.randn(T, HK, D, dtype=torch.bfloat16, device=dev) for _ in range(2))
    qr = torch.tensor([[0, 200], [200, 300]], dtype=torch.int32, device=dev)
    kr = torch.tensor([[0, 200], [50, 300]], dtype=torch.int32, device=dev)
    ty = torch.tensor([1, 0], dtype=torch.int32, device=dev)
    sc = D ** -0.5
    o_ref, m_ref = _flex_flash_attn_forward(q, k, v, None, "sh", None, None, qr, kr, ty, sc, 0.0, torch.float32)
    out = torch.zeros(T, HQ, D, dtype=torch.float32, device=dev)
    lse = torch.full((T, HQ), float("-inf"), dtype=torch.float32, device=dev)
    res = torch.ops.flex_flash_attn._flex_flash_attn_forward_compilable(
        q, k, v, None, "sh", out, lse, qr, kr, ty, sc, 0.0, None, False, False, 0, None, None, None, False, None, None, None,
        False, False, False, None, None, None, False, None, 0, False, None)
    assert res is None
    torch.testing.assert_close(out, o_ref, atol=1e-6, rtol=1e-6)
    torch.testing.assert_close(lse, m_ref.lse, atol=1e-6, rtol=1e-6)
    dq_ref, dk_ref, dv_ref, _ = _flex_flash_attn_backward(do, q, k, v, None, "sh", o_ref, m_ref.lse, None, None, None, None, qr, kr, ty,
                                                          sc, 0.0)
    dq = torch.zeros(T, HQ, D, dtype=torch.float32, device=dev)
    dk, dv = (torch.zeros(T, HK, D, dtype=torch.float32, device=dev) for _ in range(2))
    torch.ops.flex_flash_attn._flex_flash_attn_backward_compilable(
        do, q, k, v, None, "sh", out, lse, dq, dk, dv, None, qr, kr, ty, sc, 0.0, None, None, None, False, False, 0, False, None, None,
        None, False, False, False)
    # dq goes through an fp32 reduce-add whose order is not fixed: equal up to fp32 rounding; dk / dv are deterministic
    torch.testing.assert_close(dq, dq_ref, atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(dk, dk_ref, atol=1e-6, rtol=1e-6)
    torch.testing.assert_close(dv, dv_ref, atol=1e-6, rtol=1e-6)


def test_max_logits_accumulates_across_calls():
    """A given max_logits buffer is max-reduced into (the CP runtime passes one buffer through every stage)."""
    from magiattention_b200.functional.flex_flash_attn import _flex_flash_attn_forward

    dev = _dev()
    torch.manual_seed(5)
    T, H = 512, 2
    q, k, v = (torch.randn(T, H, D, dtype=torch.bfloat16, device=dev) for _ in range(3))
    mk = lambda x: torch.tensor(x, dtype=torch.int32, device=dev)
    buf = torch.full((H,), float("-inf"), device=dev)
    parts = []
    for kr in ([0, 200], [200, 512]):
        _, m = _flex_flash_attn_forward(q, k, v, None, "sh", None, None, mk([[0, T]]), mk([kr]), mk([0]), D ** -0.5, 0.0, None,
                                        return_max_logits=True, max_logits=buf)
        assert m.max_logits is buf
        parts.append(orc.max_logits(q.double(), k.double(), orc.mask_from_slices([[0, T]], [kr], [0], T, T)))
    torch.testing.assert_close(buf.double(), torch.maximum(*parts), atol=1e-4, rtol=1e-5)
    # negative maxima take the other branch of the ordered-int atomic
    buf2 = torch.full((H,), float("-inf"), device=dev)
    # (k = -q seen through a square bi-causal slice = the diagonal only: every visible logit is -|q_i|^2 * scale < 0)
    _flex_flash_attn_forward(q, -q, v, None, "sh", None, None, mk([[0, 64]]), mk([[0, 64]]), mk([3]), D ** -0.5, 0.0, None,
                             return_max_logits=True, max_logits=buf2)
    ref = orc.max_logits(q.double(), -q.double(), orc.mask_from_slices([[0, 64]], [[0, 64]], [3], T, T))
    assert (ref < 0).all()
    torch.testing.assert_close(buf2.double(), ref, atol=1e-4, rtol=1e-5)


def _random_slices(rng, T, TK, density):
    """Area-disjoint random rectangles: a random grid over (q, k) whose cells are kept with probability `density`, each
    with a random mask type — what the reference's random generator produces (test_flex_flash_attn.py:159-183)."""
    def cuts(n):
        m = int(rng.integers(1, 7))
        c = sorted(set([0, n] + [int(x) for x in rng.integers(1, max(n, 2), size=m)]))
        return [(a, b) for a, b in zip(c[:-1], c[1:]) if b > a]
    qr, kr, ty = [], [], []
    for (q0, q1) in cuts(T):
        for (k0, k1) in cuts(TK):
            if rng.random() < density:
                qr.append([q0, q1]); kr.append([k0, k1]); ty.append(int(rng.integers(0, 4)))
    return qr, kr, ty


@pytest.mark.parametrize("seed", range(6))
def test_ffa_random_slices_fuzz(seed):
    """Seeded fuzz: random area-disjoint slices of all four types over ragged (tq, tk), random GQA ratio, head_dim and
    dtype, fwd + bwd against the oracle (4 cases per seed)."""
    import numpy as np

    rng = np.random.default_rng(1234 + seed)
    for _ in range(4):
        T, TK = int(rng.integers(1, 1400)), int(rng.integers(1, 1400))
        hq, hk = [(1, 1), (2, 1), (4, 2), (6, 2), (3, 3)][int(rng.integers(0, 5))]
        d = (64, 128)[int(rng.integers(0, 2))]
        dtype = (torch.bfloat16, torch.float16)[int(rng.integers(0, 2))]
        qr, kr, ty = _random_slices(rng, T, TK, float(rng.uniform(0.3, 1.0)))
        _run_case(qr, kr, ty, T, TK, hq, hk, dtype, seed=seed, D=d)


def test_ffa_ragged_and_empty():
    # ragged sizes, an empty slice, tq not a tile multiple, tk < one block
    _run_case([[0, 77], [77, 77], [100, 333]], [[0, 50], [3, 9], [10, 50]], [0, 0, 2], 333, 50, 2, 2, torch.bfloat16)
    # no slices at all: out = 0, lse = -inf, grads = 0
    _run_case([], [], [], 130, 70, 2, 1, torch.bfloat16)


def test_ffa_golden_masks_full_size():
    """The golden-fixture masks (pinned to the reference's ref_attn_func at d=16 on CPU) re-run on the GPU
    kernels at d=128 against the same oracle."""
    files = sorted(f for f in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                   if not os.path.basename(f).startswith("sink_"))
    assert files
    for f in files[:6]:
        g = np.load(f)
        T = int(g["q"].shape[0])
        _run_case(g["q_ranges"].tolist(), g["k_ranges"].tolist(), g["types"].tolist(), T, T, 2, 1, torch.bfloat16,
                  check_bwd=False)


def test_forward_accumulate_matches_lse_merge():
    """Accumulator semantics (reference tests/test_attn/test_flex_flash_attn.py:362-540): passing (out, lse)
    in must equal correct_attn_out_lse of the separate results."""
    from magiattention_b200.functional import _flex_flash_attn_forward

    dev = _dev()
    torch.manual_seed(3)
    T, TK, H = 600, 1000, 2
    q = torch.randn(T, H, D, dtype=torch.bfloat16, device=dev)
    k = torch.randn(TK, H, D, dtype=torch.bfloat16, device=dev)
    v = torch.randn(TK, H, D, dtype=torch.bfloat16, device=dev)
    mk = lambda r: torch.tensor(r, dtype=torch.int32, device=dev).reshape(-1, 2)
    qr1, kr1, ty1 = mk([[0, 600]]), mk([[0, 400]]), torch.tensor([0], dtype=torch.int32, device=dev)
    qr2, kr2, ty2 = mk([[100, 500]]), mk([[400, 1000]]), torch.tensor([1], dtype=torch.int32, device=dev)
    scale = D ** -0.5
    o1, m1 = _flex_flash_attn_forward(q, k, v, None, "sh", None, None, qr1, kr1, ty1, scale, 0.0, torch.float32)
    o2, m2 = _flex_flash_attn_forward(q, k, v, None, "sh", None, None, qr2, kr2, ty2, scale, 0.0, torch.float32)
    ref_o, ref_l = (x.cpu() for x in orc.merge_out_lse(o1, m1.lse, o2, m2.lse))
    oa, la = o1.clone(), m1.lse.clone()
    _flex_flash_attn_forward(q, k, v, None, "sh", oa, la, qr2, kr2, ty2, scale, 0.0, torch.float32)
    torch.testing.assert_close(oa.cpu(), ref_o, atol=1e-5, rtol=1e-4)
    torch.testing.assert_close(la.cpu(), ref_l, atol=1e-5, rtol=1e-4)
    # random accumulators, as the reference test does
    o_acc = torch.randn(T, H, D, dtype=torch.float32, device=dev)
    l_acc = torch.randn(T, H, dtype=torch.float32, device=dev)
    ref_o, ref_l = (x.cpu() for x in orc.merge_out_lse(o_acc, l_acc, o2, m2.lse))
    _flex_flash_attn_forward(q, k, v, None, "sh", o_acc, l_acc, qr2, kr2, ty2, scale, 0.0, torch.float32)
    torch.testing.assert_close(o_acc.cpu(), ref_o, atol=1e-5, rtol=1e-4)
    torch.testing.assert_close(l_acc.cpu(), ref_l, atol=1e-5, rtol=1e-4)
    # and the standalone merge kernel (correct_attn_out_lse replacement)
    from magiattention_b200 import _lib
    ob, lb = o1.clone(), m1.lse.clone()
    _lib.lse_merge(ob, lb, o2, m2.lse)
    ref_o, ref_l = (x.cpu() for x in orc.merge_out_lse(o1, m1.lse, o2, m2.lse))
    torch.testing.assert_close(ob.cpu(), ref_o, atol=1e-5, rtol=1e-4)
    torch.testing.assert_close(lb.cpu(), ref_l, atol=1e-5, rtol=1e-4)


def test_backward_accumulates_dq_and_dkv():
    """bwd with given fp32 dq / dk / dv adds into them (reference :541-643); partial grads over disjoint key
    sets (global lse) sum to the full gradient."""
    from magiattention_b200.functional import _flex_flash_attn_backward, _flex_flash_attn_forward

    dev = _dev()
    torch.manual_seed(4)
    T, H = 512, 2
    q, k, v, do = (torch.randn(T, H, D, dtype=torch.bfloat16, device=dev) for _ in range(4))
    mk = lambda r: torch.tensor(r, dtype=torch.int32, device=dev).reshape(-1, 2)
    ty = torch.tensor([1], dtype=torch.int32, device=dev)
    scale = D ** -0.5
    full_q, full_k = mk([[0, T]]), mk([[0, T]])
    out, meta = _flex_flash_attn_forward(q, k, v, None, "sh", None, None, full_q, full_k, ty, scale, 0.0, None)
    dq, dk, dv, _ = _flex_flash_attn_backward(do, q, k, v, None, "sh", out, meta.lse, None, None, None, None,
                                              full_q, full_k, ty, scale, 0.0)
    # split: keys [0,200) are FULL for rows >= 200 and causal for rows < 200 ; keys [200,T) causal for rows >= 200
    qa, ka, ta = mk([[0, 200], [200, T]]), mk([[0, 200], [0, 200]]), torch.tensor([1, 0], dtype=torch.int32, device=dev)
    qb, kb, tb = mk([[200, T]]), mk([[200, T]]), torch.tensor([1], dtype=torch.int32, device=dev)
    dq2 = torch.zeros_like(dq)
    dk2 = torch.zeros_like(dk)
    dv2 = torch.zeros_like(dv)
    _flex_flash_attn_backward(do, q, k, v, None, "sh", out, meta.lse, dq2, dk2, dv2, None, qa, ka, ta, scale, 0.0)
    _flex_flash_attn_backward(do, q, k, v, None, "sh", out, meta.lse, dq2, dk2, dv2, None, qb, kb, tb, scale, 0.0)
    torch.testing.assert_close(dq2, dq, atol=2e-3, rtol=1e-3)
    torch.testing.assert_close(dk2, dk, atol=2e-3, rtol=1e-3)
    torch.testing.assert_close(dv2, dv, atol=2e-3, rtol=1e-3)
    # the same two calls sharing one BwdRowStats (the second skips the preprocess kernel): bit-identical partial sums, one
    # launch fewer, and the scratch workspace in between may be overwritten by another call
    from magiattention_b200 import _lib

    dq3, dk3, dv3 = torch.zeros_like(dq), torch.zeros_like(dk), torch.zeros_like(dv)
    rs = _lib.BwdRowStats()
    _flex_flash_attn_backward(do, q, k, v, None, "sh", out, meta.lse, dq3, dk3, dv3, None, qa, ka, ta, scale, 0.0, row_stats=rs)
    n_first = _lib.lib().b200_ffa_last_launch_count()
    assert rs.ready
    _flex_flash_attn_forward(q, k, v, None, "sh", None, None, full_q, full_k, ty, scale, 0.0, None)  # clobbers the workspace
    _flex_flash_attn_backward(do, q, k, v, None, "sh", out, meta.lse, dq3, dk3, dv3, None, qb, kb, tb, scale, 0.0, row_stats=rs)
    assert _lib.lib().b200_ffa_last_launch_count() == n_first - 1
    # dq is reduced with fp32 atomics in an unspecified order; dk / dv of disjoint calls are exact
    torch.testing.assert_close(dq3, dq2, atol=1e-5, rtol=1e-5)
    assert torch.equal(dk3, dk2) and torch.equal(dv3, dv2)
    with pytest.raises(RuntimeError, match="BwdRowStats reused"):
        _flex_flash_attn_backward(do.clone(), q, k, v, None, "sh", out, meta.lse, dq3, dk3, dv3, None, qb, kb, tb, scale, 0.0,
                                  row_stats=rs)


def test_error_behaviour():
    from magiattention_b200.functional import flex_flash_attn_func

    dev = _dev()
    q = torch.randn(64, 2, D, dtype=torch.bfloat16, device=dev)
    r = torch.tensor([[0, 64]], dtype=torch.int32, device=dev)
    with pytest.raises(AssertionError):
        flex_flash_attn_func(q, q, q, r, r, None, softcap=1.0)
    with pytest.raises(NotImplementedError):
        flex_flash_attn_func(q, q, q, r, r, None, sink=torch.zeros(1, 2, D, device=dev), sink_layout="shd")
    with pytest.raises(RuntimeError):  # more sink tokens than the kernel supports (reference: kMaxSeqlenSink = 8)
        flex_flash_attn_func(q, q, q, r, r, None, sink=torch.zeros(9, 2, device=dev))
    with pytest.raises(RuntimeError):  # sink must be fp32 [s_sink, hq]
        flex_flash_attn_func(q, q, q, r, r, None, sink=torch.zeros(1, 3, device=dev))
    with pytest.raises(RuntimeError):
        q96 = torch.randn(64, 2, 96, dtype=torch.bfloat16, device=dev)
        flex_flash_attn_func(q96, q96, q96, r, r, None)
    with pytest.raises(RuntimeError):
        flex_flash_attn_func(q.cpu(), q.cpu(), q.cpu(), r.cpu(), r.cpu(), None)


@pytest.mark.slow
def test_full_size_properties_c2_like():
    """BASELINE config-2-like size (64K packed varlen, causal + sliding window): size-independent properties —
    linearity in V, lse invariance to V, and row-block agreement with the oracle on a sampled document."""
    from magiattention_b200.functional import flex_flash_attn_func

    dev = _dev()
    torch.manual_seed(5)
    T, H = 65536, 4
    lens = [4096, 12288, 1024, 16384, 8192, 3000, 5192, 15360]
    assert sum(lens) == T
    qr, kr, ty = [], [], []
    pos = 0
    for i, n in enumerate(lens):
        if i % 2 == 0:
            qr.append([pos, pos + n]); kr.append([pos, pos + n]); ty.append(1)
        else:
            w = 1024
            qr.append([pos, pos + w]); kr.append([pos, pos + w]); ty.append(1)
            qr.append([pos + w, pos + n]); kr.append([pos, pos + n]); ty.append(3)
        pos += n
    mk = lambda r: torch.tensor(r, dtype=torch.int32, device=dev).reshape(-1, 2)
    q, k, v, v2 = (torch.randn(T, H, D, dtype=torch.bfloat16, device=dev) for _ in range(4))
    qr_t, kr_t, ty_t = mk(qr), mk(kr), torch.tensor(ty, dtype=torch.int32, device=dev)
    o1, m1 = flex_flash_attn_func(q, k, v, qr_t, kr_t, ty_t)
    o2, m2 = flex_flash_attn_func(q, k, v2, qr_t, kr_t, ty_t)
    o3, m3 = flex_flash_attn_func(q, k, (v.float() + v2.float()).to(torch.bfloat16), qr_t, kr_t, ty_t)
    assert torch.equal(m1.lse, m2.lse)
    torch.testing.assert_close(o3.float(), o1.float() + o2.float(), atol=6e-2, rtol=5e-2)
    # oracle on document 2 (1024 rows, causal) and the head of document 3's band
    for (a, b) in ((4096 + 12288, 4096 + 12288 + 1024),):
        sub = slice(a, b)
        mask = orc.mask_from_slices([[0, b - a]], [[0, b - a]], [1], b - a, b - a)
        ro, rl = orc.attn_fwd(q[sub], k[sub], v[sub], mask, high_precision=True)
        lo, ll = orc.attn_fwd(q[sub], k[sub], v[sub], mask, high_precision=False)
        orc.check_close("out", o1[sub], ro, lo, "out")
        orc.check_close("lse", m1.lse[sub], rl, ll, "lse")


def test_plan_overflow_falls_back_to_slice_scan():
    """With a workspace too small for the tile->slice CSR the kernels must scan all slices instead (same results)."""
    import ctypes as C
    from magiattention_b200 import _lib

    dev = _dev()
    torch.manual_seed(9)
    cfg = next(c for c in mask_configs(1) if c["name"] == "sparse_attn_2k")
    T, H = cfg["seqlen"], 2
    q, k, v = (torch.randn(T, H, D, dtype=torch.bfloat16, device=dev) for _ in range(3))
    mk = lambda r: torch.tensor(r, dtype=torch.int32, device=dev).reshape(-1, 2)
    qr, kr, ty = mk(cfg["q_ranges"]), mk(cfg["k_ranges"]), torch.tensor(cfg["types"], dtype=torch.int32, device=dev)
    out_ref = torch.empty_like(q)
    lse_ref = torch.empty(T, H, dtype=torch.float32, device=dev)
    _lib.ffa_fwd(q, k, v, out_ref, lse_ref, qr, kr, ty, D ** -0.5)
    # same call through the raw ABI with a workspace that only fits the CSR header + offsets
    p = _lib.FwdParams()
    out = torch.empty_like(q)
    lse = torch.empty(T, H, dtype=torch.float32, device=dev)
    p.q, p.k, p.v, p.out, p.lse = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr()
    p.q_ranges, p.k_ranges, p.attn_type_map, p.num_slices = qr.data_ptr(), kr.data_ptr(), ty.data_ptr(), qr.shape[0]
    p.total_q = p.total_k = T
    p.num_heads_q = p.num_heads_kv = H
    p.head_dim = D
    for n, t in (("q", q), ("k", k), ("v", v), ("o", out)):
        setattr(p, f"{n}_stride_t", t.stride(0)); setattr(p, f"{n}_stride_h", t.stride(1))
    p.lse_stride_t, p.lse_stride_h = lse.stride(0), lse.stride(1)
    p.softmax_scale, p.dtype, p.out_dtype = D ** -0.5, _lib.B200_BF16, _lib.B200_BF16
    ws = torch.empty(320, dtype=torch.uint8, device=dev)  # header + offsets + room for only 16 CSR entries
    assert ws.numel() < _lib.lib().b200_ffa_fwd_workspace_bytes(C.byref(p))
    p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
    rc = _lib.lib().b200_ffa_fwd(C.byref(p), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, _lib.lib().b200_ffa_last_error()
    torch.cuda.synchronize()
    assert torch.equal(out, out_ref) and torch.equal(lse, lse_ref)


def test_strided_inputs_packed_qkv():
    """q/k/v as views into one packed [T, 3, H, D] tensor (token stride 3*H*D): the TMA maps carry the strides."""
    from magiattention_b200 import _lib

    dev = _dev()
    torch.manual_seed(10)
    T, H = 777, 2
    qkv = torch.randn(T, 3, H, D, dtype=torch.bfloat16, device=dev)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    r = torch.tensor([[0, T]], dtype=torch.int32, device=dev)
    ty = torch.tensor([1], dtype=torch.int32, device=dev)
    out = torch.empty(T, H, D, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(T, H, dtype=torch.float32, device=dev)
    _lib.ffa_fwd(q, k, v, out, lse, r, r, ty, D ** -0.5)
    out2 = torch.empty_like(out)
    lse2 = torch.empty_like(lse)
    _lib.ffa_fwd(q.contiguous(), k.contiguous(), v.contiguous(), out2, lse2, r, r, ty, D ** -0.5)
    assert torch.equal(out, out2) and torch.equal(lse, lse2)


def test_baseline_config_c1():
    """BASELINE.json configs[0]: single-GPU fwd, seq=4096, nheads=16, d=128, full-causal, bf16 — area 8 390 656."""
    qr = kr = [[0, 4096]]
    assert orc.mask_area(qr, kr, [1]) == 8390656
    _run_case(qr, kr, [1], 4096, 4096, 16, 16, torch.bfloat16, check_bwd=False)


def test_block_sparse_gqa_c5_like():
    """C5-like (scaled): GQA 8:1, video-gen style varlen block-causal rectangles — each q block attends its own
    block, the previous two blocks and the global first block — as FULL slices; fwd + bwd."""
    torch.manual_seed(12)
    edges = [0, 300, 812, 1324, 1580, 2604, 3116, 3500, 4096]
    qr, kr = [], []
    for i in range(len(edges) - 1):
        q = [edges[i], edges[i + 1]]
        lo = max(i - 2, 0)
        qr.append(q); kr.append([edges[lo], edges[i + 1]])
        if lo > 0:
            qr.append(q); kr.append([edges[0], edges[1]])
    _run_case(qr, kr, [0] * len(qr), 4096, 4096, 8, 1, torch.bfloat16)


def test_range_reduce_sum_multi_source():
    """group_reduce's receive-side kernel: several sources folding into the SAME destination span (cp >= 3) must be
    race-free and exact in fp32 (regression: independent blocks did read-modify-write on the same rows)."""
    from magiattention_b200 import _lib

    dev = _dev()
    torch.manual_seed(13)
    rows, feat = 4096, 256
    dst = torch.randn(rows, feat, dtype=torch.float32, device=dev)
    ref = dst.double().clone()
    src = torch.randn(8192, feat, dtype=torch.float32, device=dev).to(torch.bfloat16)
    d_spans, s_spans, spos = [], [], 0
    for (a, b, nsrc) in ((0, 1000, 3), (1000, 1024, 1), (2048, 4096, 2)):
        for _ in range(nsrc):
            d_spans.append([a, b]); s_spans.append([spos, spos + (b - a)])
            ref[a:b] += src[spos:spos + (b - a)].double()
            spos += b - a
    mk = lambda r: torch.tensor(r, dtype=torch.int32, device=dev)
    for _ in range(3):  # repeat: a race would show up as run-to-run differences
        out = dst.clone()
        _lib.range_reduce_sum(out, src, mk(d_spans), mk(s_spans))
        torch.testing.assert_close(out.double(), ref, atol=1e-5, rtol=1e-6)
