"""Generate tests/golden/functools_masks.json by running the REFERENCE's mask helpers
(magi_attention/api/functools.py: infer_attn_mask_from_sliding_window :180-332, infer_attn_mask_from_cu_seqlens :335-420)
on a sweep of inputs.  Runs only in the build container (needs /root/reference); the JSON travels.

    python tests/golden/gen_functools_golden.py
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from gen_golden import _import_reference  # noqa: E402  (same stubs)


def main():
    _import_reference()
    from magi_attention.api.functools import infer_attn_mask_from_cu_seqlens, infer_attn_mask_from_sliding_window
    from magi_attention.common.range import AttnRange

    def pack(qr, kr, ty):
        return {"q_ranges": qr.to_naive_ranges(), "k_ranges": kr.to_naive_ranges(), "types": [t.to_int_type() for t in ty]}

    sw = []
    for (q0, q1, k0, k1) in [(5, 15, 5, 15), (0, 12, 0, 12), (0, 6, 0, 14), (3, 17, 2, 9), (0, 1, 0, 1), (0, 9, 4, 13), (2, 10, 0, 30)]:
        for w in [(2, 3), (0, 0), (-1, -1), (-1, 0), (0, -1), (1, -1), (-1, 2), (3, 0), (0, 4), (5, 5), (100, 1), (1, 100), (7, 2)]:
            qr, kr, ty = infer_attn_mask_from_sliding_window(AttnRange(q0, q1), AttnRange(k0, k1), w)
            sw.append({"q_range": [q0, q1], "k_range": [k0, k1], "window": list(w), **pack(qr, kr, ty)})
    cu = []
    for cq, ck in [([0, 7, 19, 32], [0, 7, 19, 32]), ([0, 5, 5, 20], [0, 5, 5, 20]), ([0, 10, 24], [0, 16, 40]), ([0, 16, 40], [0, 10, 24]),
                   ([0, 64], [0, 64])]:
        for causal, w in [(False, (-1, -1)), (True, (-1, -1)), (False, (3, 0)), (False, (2, 2)), (False, (-1, 0)), (False, (0, -1)),
                          (False, (4, 1)), (False, (30, 0))]:
            qr, kr, ty, tq, tk = infer_attn_mask_from_cu_seqlens(torch.tensor(cq, dtype=torch.int32), torch.tensor(ck, dtype=torch.int32),
                                                                 causal=causal, window_size=w)
            cu.append({"cu_q": cq, "cu_k": ck, "causal": causal, "window": list(w), "total_q": int(tq), "total_k": int(tk), **pack(qr, kr, ty)})
    path = os.path.join(HERE, "functools_masks.json")
    json.dump({"sliding_window": sw, "cu_seqlens": cu}, open(path, "w"))
    print(f"{len(sw)} sliding-window cases, {len(cu)} cu_seqlens cases -> {os.path.getsize(path)} B")


if __name__ == "__main__":
    main()
