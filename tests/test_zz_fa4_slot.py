"""The reference runtime's Blackwell kernel slot (`fa4_fwd` / `fa4_bwd`, reference functional/fa4.py:125-321) served by this
backend.  CPU: the glue binds every argument to the right parameter of the raw launchers.  GPU (this file sorts last on
purpose): the slot gives the same out / lse / dq / dk / dv as `flex_flash_attn_func` + autograd on the same slices."""
import inspect

import pytest
import torch

D = 128


class _Arg:
    """Minimal stand-in for the reference's AttnArg (calc_meta.py:146-159): to_ffa_args(is_bwd) -> dict of int32 tensors."""

    def __init__(self, qr, kr, ty, device="cpu"):
        mk = lambda x, shape: torch.tensor(x, dtype=torch.int32, device=device).reshape(shape)
        self.d = dict(q_ranges=mk(qr, (-1, 2)), k_ranges=mk(kr, (-1, 2)), attn_type_map=mk(ty, (-1,)))
        self.calls = []

    def to_ffa_args(self, is_bwd=False):
        self.calls.append(is_bwd)
        return self.d


def test_slot_glue_binds_launcher_parameters(monkeypatch):
    from magiattention_b200.common.forward_meta import AttnForwardMeta
    from magiattention_b200.functional import fa4
    from magiattention_b200.functional import flex_flash_attn as ffa

    seen = {}

    def fake_fwd(*a, **kw):
        b = inspect.signature(ffa._flex_flash_attn_forward).bind(*a, **kw)
        seen["fwd"] = b.arguments
        q = b.arguments["q"]
        return torch.zeros_like(q), AttnForwardMeta(lse=torch.zeros(q.shape[0], q.shape[1]), max_logits=None)

    def fake_bwd(*a, **kw):
        b = inspect.signature(ffa._flex_flash_attn_backward).bind(*a, **kw)
        seen["bwd"] = b.arguments
        q, k, v = (b.arguments[n] for n in ("q", "k", "v"))
        dk_type = b.arguments.get("dk_type")
        return torch.zeros_like(q, dtype=torch.float32), torch.zeros_like(k, dtype=dk_type), torch.zeros_like(v, dtype=dk_type), None

    monkeypatch.setattr(fa4, "_flex_flash_attn_forward", fake_fwd)
    monkeypatch.setattr(fa4, "_flex_flash_attn_backward", fake_bwd)
    q, k, v, do = (torch.randn(64, 2, D, dtype=torch.bfloat16) for _ in range(4))
    arg = _Arg([[0, 64]], [[0, 64]], [1])
    out, lse = fa4.fa4_fwd(q, k, v, None, arg, softmax_scale=0.25)
    f = seen["fwd"]
    assert f["q"] is q and f["k"] is k and f["v"] is v and f["sink"] is None and f["sink_layout"] == "sh"
    assert f["out"] is None and f["lse"] is None  # the slot starts a fresh result (no accumulator)
    assert f["q_ranges"] is arg.d["q_ranges"] and f["k_ranges"] is arg.d["k_ranges"] and f["attn_type_map"] is arg.d["attn_type_map"]
    assert f["softmax_scale"] == 0.25 and f["softcap"] == 0.0 and f["out_type"] == torch.bfloat16
    assert out.dtype == torch.bfloat16 and lse.shape == (64, 2) and arg.calls == [False]
    fa4.fa4_fwd(q, k, v, None, arg)
    assert seen["fwd"]["softmax_scale"] == D ** -0.5  # default scale (reference fa4.py:139-140)
    dq, dk, dv, dsink = fa4.fa4_bwd(do, q, k, v, None, out, lse, arg, softmax_scale=0.25)
    b = seen["bwd"]
    assert b["dout"] is do and b["q"] is q and b["k"] is k and b["v"] is v and b["out"] is out and b["lse"] is lse
    assert b["dq"] is None and b["dk"] is None and b["dv"] is None and b["dsink"] is None
    assert b["q_ranges"] is arg.d["q_ranges"] and b["softmax_scale"] == 0.25 and b["softcap"] == 0.0
    assert b["dk_type"] == torch.bfloat16 and b["dv_type"] == torch.bfloat16 and b["deterministic"] is False
    assert dq.dtype == torch.bfloat16 and dk.dtype == torch.bfloat16 and dsink is None and arg.calls[-1] is True
    # the reference's slot signatures, name for name
    assert list(inspect.signature(fa4.fa4_fwd).parameters) == ["q", "k", "v", "sink", "attn_arg", "softmax_scale", "softcap", "sink_layout"]
    assert list(inspect.signature(fa4.fa4_bwd).parameters) == ["do", "q", "k", "v", "sink", "o", "lse", "attn_arg", "softmax_scale",
                                                              "softcap", "sink_layout", "deterministic"]


@pytest.mark.gpu
@pytest.mark.parametrize("with_sink", [False, True])
def test_slot_equals_flex_flash_attn_func(with_sink):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from magiattention_b200.functional import fa4, flex_flash_attn_func
    from magiattention_b200.meta.planner import AttnArg

    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    T, HQ, HK = 1024, 4, 2
    q, do = (torch.randn(T, HQ, D, dtype=torch.bfloat16, device=dev) for _ in range(2))
    k, v = (torch.randn(T, HK, D, dtype=torch.bfloat16, device=dev) for _ in range(2))
    sink = torch.randn(2, HQ, dtype=torch.float32, device=dev) if with_sink else None
    qr, kr, ty = [[0, 384], [384, 1024], [384, 1024]], [[0, 384], [0, 256], [384, 1024]], [1, 0, 1]
    q_, k_, v_ = (x.clone().requires_grad_(True) for x in (q, k, v))
    s_ = sink.clone().requires_grad_(True) if with_sink else None
    mk = lambda x, shape: torch.tensor(x, dtype=torch.int32, device=dev).reshape(shape)
    out_ref, meta = flex_flash_attn_func(q_, k_, v_, mk(qr, (-1, 2)), mk(kr, (-1, 2)), mk(ty, (-1,)), sink=s_)
    out_ref.backward(do)
    for arg in (AttnArg([tuple(x) for x in qr], [tuple(x) for x in kr], ty), _Arg(qr, kr, ty, dev)):
        out, lse = fa4.fa4_fwd(q, k, v, sink, arg)
        assert torch.equal(out, out_ref.detach()) and torch.equal(lse, meta.lse)
        dq, dk, dv, dsink = fa4.fa4_bwd(do, q, k, v, sink, out, lse, arg)
        torch.testing.assert_close(dq.float(), q_.grad.float(), atol=2e-2, rtol=2e-2)  # fp32 atomics order + one bf16 rounding
        assert torch.equal(dk, k_.grad) and torch.equal(dv, v_.grad)
        if with_sink:
            torch.testing.assert_close(dsink, s_.grad.float(), atol=1e-4, rtol=1e-4)
        else:
            assert dsink is None
