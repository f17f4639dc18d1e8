"""Generate tests/golden/dispatch_partitions.json: the chunk -> rank partitions the REFERENCE's dispatch planner
(magi_attention/meta/_make_dispatch_meta.py:56 make_dispatch_meta_from_qk_ranges, default MinHeapDispatchAlg) emits for a set of
masks, used to pin the load balance of our planner.  Build container only (needs /root/reference).

    python tests/golden/gen_dispatch_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from gen_golden import _import_reference  # noqa: E402
from mask_configs import mask_configs  # noqa: E402


def main():
    _import_reference()
    from magi_attention.common.enum import AttnMaskType
    from magi_attention.common.ranges import AttnRanges
    from magi_attention.meta._make_dispatch_meta import make_dispatch_meta_from_qk_ranges
    from magi_attention.meta.solver.dispatch_solver import DispatchConfig, MinHeapDispatchAlg

    cases = []
    cfgs = {c["name"]: c for c in mask_configs(1)}
    todo = [("causal_64chunks", [[0, 4096]], [[0, 4096]], [1], 4096, 64)]
    for name in ("varlen_causal_1k", "sliding_window_causal", "all_types_aspects", "varlen_block_causal_2k", "sparse_attn_2k"):
        c = cfgs[name]
        todo.append((name, c["q_ranges"], c["k_ranges"], c["types"], c["seqlen"], 64))
    for name, qr, kr, ty, T, chunk in todo:
        for cp in (2, 4, 8):
            pad = (-T) % (chunk * cp)
            total = T + pad
            meta_q, _ = make_dispatch_meta_from_qk_ranges(
                q_ranges=AttnRanges.from_ranges(qr), k_ranges=AttnRanges.from_ranges(kr),
                attn_mask_type=[AttnMaskType.from_int_type(t) for t in ty], total_seqlen_q=total, total_seqlen_k=total,
                chunk_size=chunk, cp_size=cp, cp_rank=0, dispatch_config=DispatchConfig(alg=MinHeapDispatchAlg()),
                is_same_source=True, is_q_permutable=True, is_k_permutable=True)
            cases.append({"name": name, "q_ranges": qr, "k_ranges": kr, "types": ty, "total_seqlen": total, "chunk_size": chunk,
                          "cp_size": cp, "partitions": [list(map(int, p)) for p in meta_q.partitions]})
            print(name, cp, [p[:4] for p in meta_q.partitions][:2])
    path = os.path.join(HERE, "dispatch_partitions.json")
    json.dump(cases, open(path, "w"))
    print(f"{len(cases)} cases -> {os.path.getsize(path)} B")


if __name__ == "__main__":
    main()
