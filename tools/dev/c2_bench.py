"""BASELINE config C2-like measurement: one B200, packed varlen batch of 64K tokens (mixed document lengths, half the
documents causal, half sliding-window(1024)-causal), 32/32 heads, d=128, bf16, fwd and bwd TFLOP/s."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402

from magiattention_b200 import _lib  # noqa: E402
from magiattention_b200.api import infer_attn_mask_from_cu_seqlens  # noqa: E402
from oracle import ffa_oracle as orc  # noqa: E402  (area / FLOP accounting only)


def c5_like():
    """BASELINE config C5-like on one GPU: 128K tokens (one rank's share of the 1M sequence), 64:8 GQA, block-sparse
    video-style mask: variable blocks of 512..4096 tokens, each q block sees itself, the two previous blocks and the first
    (global) block, as FULL rectangles."""
    dev = torch.device("cuda:0")
    torch.manual_seed(42)
    T, HQ, HK, D = 131072, 64, 8, 128
    g = torch.Generator().manual_seed(42)
    cuts = [0]
    while cuts[-1] < T:
        cuts.append(min(T, cuts[-1] + 512 * int(torch.randint(1, 9, (1,), generator=g))))
    blocks = list(zip(cuts[:-1], cuts[1:]))
    qr, kr = [], []
    for i, (a, b) in enumerate(blocks):
        for j in sorted({0, i - 2, i - 1, i}):
            if j >= 0:
                qr.append([a, b]); kr.append(list(blocks[j]))
    ty = [0] * len(qr)
    area = orc.mask_area(qr, kr, ty)
    print(f"C5-like: {len(blocks)} blocks, {len(ty)} slices, area {area} ({area / (T * T):.4f} dense)")
    q, do = (torch.randn(T, HQ, D, dtype=torch.bfloat16, device=dev) for _ in range(2))
    k, v = (torch.randn(T, HK, D, dtype=torch.bfloat16, device=dev) for _ in range(2))
    _time(q, k, v, do, qr, kr, ty, area, "C5-like")


def _time(q, k, v, do, qr, kr, ty, area, tag):
    dev = q.device
    T, H, D = q.shape
    mk = lambda r: torch.tensor(r, dtype=torch.int32, device=dev).reshape(-1, 2)
    qr_t, kr_t, ty_t = mk(qr), mk(kr), torch.tensor(ty, dtype=torch.int32, device=dev)
    out = torch.empty_like(q)
    lse = torch.empty(T, H, dtype=torch.float32, device=dev)
    dq = torch.zeros(T, H, D, dtype=torch.float32, device=dev)
    dk, dv = torch.empty_like(k), torch.empty_like(v)
    sc = D ** -0.5
    f = lambda: _lib.ffa_fwd(q, k, v, out, lse, qr_t, kr_t, ty_t, sc)
    b = lambda: _lib.ffa_bwd(do, q, k, v, out, lse, dq, dk, dv, qr_t, kr_t, ty_t, sc)
    for name, fn, mult in (("fwd", f, 1.0), ("bwd", b, 2.5)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"[PERF] {tag} {name}: {ms:8.3f} ms  {mult * 4.0 * area * H * D / ms / 1e9:8.1f} TFLOP/s", flush=True)


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(42)
    T, H, D = 65536, 32, 128
    g = torch.Generator().manual_seed(42)
    lens = []
    while sum(lens) < T:
        # long-tailed mix: many short documents, a few long ones, capped at T/4
        n = int(min(T // 4, max(64, torch.distributions.LogNormal(7.5, 1.2).sample().item())))
        lens.append(min(n, T - sum(lens)))
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    qr, kr, ty = [], [], []
    for i in range(len(lens)):
        a, b, c, _, _ = infer_attn_mask_from_cu_seqlens([cu[i], cu[i + 1]], [cu[i], cu[i + 1]], causal=(i % 2 == 0),
                                                        window_size=(-1, -1) if i % 2 == 0 else (1024, 0))
        qr += a.to_naive_ranges(); kr += b.to_naive_ranges(); ty += [t.to_int_type() for t in c]
    area = orc.mask_area(qr, kr, ty)
    print(f"C2-like: {len(lens)} docs (min {min(lens)}, max {max(lens)}), {len(ty)} slices, area {area} "
          f"({area / (T * (T + 1) / 2):.3f} of full causal)")
    q, k, v, do = (torch.randn(T, H, D, dtype=torch.bfloat16, device=dev) for _ in range(4))
    _time(q, k, v, do, qr, kr, ty, area, "C2-like")


if __name__ == "__main__":
    c5_like() if "--c5" in sys.argv else main()
