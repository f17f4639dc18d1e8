"""Write (or --check) tests/golden/api_signatures.json + api_key_fields.json from the REFERENCE's own modules, imported on
CPU with the SURVEY App. C stubs (debugpy stub, empty magi_attn_ext / magi_attn_comm modules)."""
import dataclasses
import inspect
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path[:0] = [REF, os.path.join(HERE, "..", "..", "baseline", "stubs")]
for m in ("magi_attention.magi_attn_ext", "magi_attention.magi_attn_comm"):
    sys.modules.setdefault(m, types.ModuleType(m))
import warnings  # noqa: E402

warnings.filterwarnings("ignore")
import magi_attention  # noqa: E402,F401
import magi_attention.dist_attn_runtime_mgr as mgr  # noqa: E402
from magi_attention.api import magi_attn_interface as api  # noqa: E402
from magi_attention.comm.primitive.grpcoll import _group_collective as gc  # noqa: E402
from magi_attention.functional import fa4, flex_flash_attn as ffa  # noqa: E402


def sig(fn):
    fn = getattr(fn, "__wrapped__", fn)
    out = []
    for p in inspect.signature(fn).parameters.values():
        out.append([p.name, p.kind.name, None if p.default is inspect.Parameter.empty else repr(p.default)])
    return out


M = mgr.DistAttnRuntimeMgr
table = {
    "init_dist_attn_runtime_key": mgr.init_dist_attn_runtime_key, "init_dist_attn_runtime_mgr": mgr.init_dist_attn_runtime_mgr,
    "DistAttnRuntimeMgr.__init__": M.__init__, "DistAttnRuntimeMgr.dispatch_qo": M.dispatch_qo,
    "DistAttnRuntimeMgr.dispatch_kv": M.dispatch_kv, "DistAttnRuntimeMgr.undispatch_qo": M.undispatch_qo,
    "DistAttnRuntimeMgr.undispatch_kv": M.undispatch_kv, "DistAttnRuntimeMgr.calc_attn": M.calc_attn,
    "DistAttnRuntimeMgr.get_position_ids": M.get_position_ids, "DistAttnRuntimeMgr.roll": M.roll,
    "fa4_fwd": fa4.fa4_fwd, "fa4_bwd": fa4.fa4_bwd, "ffa_fa4_func": fa4.ffa_fa4_func,
    "flex_flash_attn_func": ffa.flex_flash_attn_func,
    "magi_attn_flex_key": api.magi_attn_flex_key, "magi_attn_varlen_key": api.magi_attn_varlen_key,
    "dispatch": api.dispatch, "undispatch": api.undispatch, "calc_attn": api.calc_attn, "get_position_ids": api.get_position_ids,
    "roll": api.roll, "make_flex_key_for_new_mask_after_dispatch": api.make_flex_key_for_new_mask_after_dispatch,
    "make_varlen_key_for_new_mask_after_dispatch": api.make_varlen_key_for_new_mask_after_dispatch,
    "group_cast": gc.group_cast, "group_reduce": gc.group_reduce,
}
sigs = {k: sig(v) for k, v in table.items()}
fields = [f.name for f in dataclasses.fields(mgr.DistAttnRuntimeKey)]
p1, p2 = os.path.join(HERE, "api_signatures.json"), os.path.join(HERE, "api_key_fields.json")
if "--check" in sys.argv:
    ok = json.load(open(p1)) == sigs and json.load(open(p2)) == fields
    print("pinned signatures match the reference" if ok else "PINNED SIGNATURES ARE STALE")
    sys.exit(0 if ok else 1)
json.dump(sigs, open(p1, "w"), indent=1)
json.dump(fields, open(p2, "w"), indent=1)
print("wrote", p1, p2)
