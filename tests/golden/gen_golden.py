"""Generate tests/golden/*.npz by running the REFERENCE's own CPU oracle.

Runs only in the build container (needs /root/reference, read-only): imports
``magi_attention.testing.ref_attn.ref_attn_func`` (testing/ref_attn.py:638-749) and
``magi_attention.utils.make_attn_mask_from_ffa_args`` (utils/_utils.py:880-922) with the three
stubs described in SURVEY.md App. C, evaluates every mask config of tests/mask_configs.py at
scale 4 in fp64 (fwd + autograd bwd) and in bf16 (fwd), and stores inputs + outputs.
The committed fixtures travel to the GPU box; this script does not need to.

    python tests/golden/gen_golden.py
"""
from __future__ import annotations

import importlib.util
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from mask_configs import mask_configs  # noqa: E402

REF = "/root/reference"


def _import_reference():
    stub_dir = tempfile.mkdtemp(prefix="stubs_")
    os.makedirs(os.path.join(stub_dir, "debugpy"))
    with open(os.path.join(stub_dir, "debugpy", "__init__.py"), "w") as f:
        f.write("def listen(*a, **k):\n    pass\n\ndef wait_for_client(*a, **k):\n    pass\n")
    sys.path[:0] = [stub_dir, REF]
    for n in ("magi_attention.magi_attn_ext", "magi_attention.magi_attn_comm"):
        m = types.ModuleType(n)
        m.__file__ = "/tmp/stub.so"
        sys.modules[n] = m
    import magi_attention  # noqa: F401
    from magi_attention.common.ranges import AttnRanges
    from magi_attention.utils import make_attn_mask_from_ffa_args

    spec = importlib.util.spec_from_file_location(
        "ref_attn_by_path", os.path.join(REF, "magi_attention/testing/ref_attn.py")
    )
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return AttnRanges, make_attn_mask_from_ffa_args, mod.ref_attn_func


def main():
    AttnRanges, make_mask, ref_attn_func = _import_reference()
    HQ, HK, D = 2, 1, 16
    for idx, cfg in enumerate(mask_configs(scale=4)):
        torch.manual_seed(1000 + idx)
        T = cfg["seqlen"]
        q = torch.randn(T, HQ, D).to(torch.bfloat16)
        k = torch.randn(T, HK, D).to(torch.bfloat16)
        v = torch.randn(T, HK, D).to(torch.bfloat16)
        do = torch.randn(T, HQ, D).to(torch.bfloat16)
        mask = make_mask(
            q_ranges=AttnRanges.from_ranges(cfg["q_ranges"]),
            k_ranges=AttnRanges.from_ranges(cfg["k_ranges"]),
            attn_type_map=cfg["types"],
            total_seqlen_q=T,
            total_seqlen_k=T,
            device="cpu",
        )
        q64, k64, v64 = (x.double().requires_grad_(True) for x in (q, k, v))
        out, meta = ref_attn_func(q64, k64, v64, mask, backend="torch", high_precision=True, return_lse=True)
        out.backward(do.double())
        out_lp, meta_lp = ref_attn_func(q, k, v, mask, backend="torch", high_precision=False, return_lse=True)
        path = os.path.join(HERE, f"{cfg['name']}.npz")
        np.savez_compressed(
            path,
            q=q.view(torch.int16).numpy(), k=k.view(torch.int16).numpy(), v=v.view(torch.int16).numpy(),
            do=do.view(torch.int16).numpy(),
            q_ranges=np.array(cfg["q_ranges"], dtype=np.int32), k_ranges=np.array(cfg["k_ranges"], dtype=np.int32),
            types=np.array(cfg["types"], dtype=np.int32),
            mask_area=np.int64(int(mask.sum())),
            out=out.detach().numpy(), lse=meta.lse.detach().double().numpy(),
            dq=q64.grad.numpy(), dk=k64.grad.numpy(), dv=v64.grad.numpy(),
            out_lp=out_lp.detach().float().numpy(), lse_lp=meta_lp.lse.detach().float().numpy(),
        )
        print(f"{cfg['name']:32s} T={T:5d} slices={len(cfg['types']):3d} area={int(mask.sum())} -> {os.path.getsize(path)} B")


SINK_CASES = (("causal", "sh", 1), ("all_types_aspects", "sh", 4), ("uncovered_rows", "sh", 8),
              ("varlen_causal_1k", "ssh", 2), ("sliding_window_causal", "ssh", 6))


def main_sink():
    """Attention sink + max_logits fixtures (reference: ref_attn_func(sink=..., sink_layout=..., return_max_logits=True),
    testing/ref_attn.py:638-749; dsink through the reference's own autograd function)."""
    AttnRanges, make_mask, ref_attn_func = _import_reference()
    HQ, HK, D = 2, 1, 16
    cfgs = {c["name"]: c for c in mask_configs(scale=4)}
    for idx, (name, layout, s_sink) in enumerate(SINK_CASES):
        cfg = cfgs[name]
        torch.manual_seed(2000 + idx)
        T = cfg["seqlen"]
        q = torch.randn(T, HQ, D).to(torch.bfloat16)
        k = torch.randn(T, HK, D).to(torch.bfloat16)
        v = torch.randn(T, HK, D).to(torch.bfloat16)
        do = torch.randn(T, HQ, D).to(torch.bfloat16)
        sink = torch.randn((s_sink, HQ) if layout == "sh" else (T, s_sink, HQ), dtype=torch.float32)
        mask = make_mask(
            q_ranges=AttnRanges.from_ranges(cfg["q_ranges"]), k_ranges=AttnRanges.from_ranges(cfg["k_ranges"]),
            attn_type_map=cfg["types"], total_seqlen_q=T, total_seqlen_k=T, device="cpu",
        )
        q64, k64, v64 = (x.double().requires_grad_(True) for x in (q, k, v))
        s64 = sink.double().requires_grad_(True)
        out, meta = ref_attn_func(q64, k64, v64, mask, sink=s64, sink_layout=layout, backend="torch", high_precision=True,
                                  return_lse=True, return_max_logits=True)
        out.backward(do.double())
        path = os.path.join(HERE, f"sink_{layout}_{name}.npz")
        np.savez_compressed(
            path,
            q=q.view(torch.int16).numpy(), k=k.view(torch.int16).numpy(), v=v.view(torch.int16).numpy(),
            do=do.view(torch.int16).numpy(), sink=sink.numpy(), sink_layout=np.array(layout),
            q_ranges=np.array(cfg["q_ranges"], dtype=np.int32), k_ranges=np.array(cfg["k_ranges"], dtype=np.int32),
            types=np.array(cfg["types"], dtype=np.int32),
            out=out.detach().numpy(), lse=meta.lse.detach().double().numpy(), max_logits=meta.max_logits.detach().double().numpy(),
            dq=q64.grad.numpy(), dk=k64.grad.numpy(), dv=v64.grad.numpy(), dsink=s64.grad.numpy(),
        )
        print(f"sink {layout:3s} {name:28s} T={T:5d} s_sink={s_sink} -> {os.path.getsize(path)} B")


if __name__ == "__main__":
    if "--sink-only" not in sys.argv:
        main()
    main_sink()
