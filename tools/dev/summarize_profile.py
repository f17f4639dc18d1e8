"""Turn gpurun_out/{launches.csv,prof.ncu-rep,bench_n1.json} into the tracked summaries under profiles/."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(ROOT, "profiles")
G = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

# ---- launch list
p = os.path.join(G, "launches.csv")
if os.path.exists(p):
    rows = list(csv.reader(l for l in open(p) if not l.startswith("==")))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except Exception:
            continue
        v = {"ns": v / 1e6, "us": v / 1e3, "usecond": v / 1e3, "ms": v, "msecond": v}.get(r[ui], v / 1e6)
        agg.setdefault(r[ki], [0, 0.0])
        agg[r[ki]][0] += 1
        agg[r[ki]][1] += v
    tot = sum(v for _, v in agg.values())
    with open(os.path.join(OUT, f"{tag}_launches.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 1 --warmup 1 "
                "--no-e2e --no-cpu --seqlen 65536\n# (cold-cache, serialised: compare SHARES)\n")
        f.write(f"{'ms':>10s} {'share':>7s} {'count':>6s}  kernel\n")
        for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{v:10.3f} {100 * v / tot:6.2f}% {c:6d}  {n[:110]}\n")
        f.write(f"total {tot:.3f} ms\n")

# ---- full capture
rep = os.path.join(G, "prof.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    keys = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size"]
    with open(os.path.join(OUT, f"{tag}_ncu_full.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:ffa_(fwd|bwd)_kernel -c 2  python bench.py "
                "--steps 1 --warmup 1 --no-e2e --no-cpu --seqlen 32768   (32 heads, causal, bf16)\n")
        for r in rows[2:]:
            f.write(f"\n== {r[idx['Kernel Name']][:100]}\n")
            for k in keys:
                if k in idx:
                    f.write(f"   {k:78s} {r[idx[k]]:>18s} {units[idx[k]]}\n")

# ---- DRAM traffic of the attention kernels at the full bench size (one-pass ncu metrics run)
t = os.path.join(G, "traffic_256k.csv")
if os.path.exists(t):
    rows = list(csv.reader(l for l in open(t) if not l.startswith("==")))
    if rows:
        hdr = rows[0]
        ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        out = {}
        for r in rows[1:]:
            try:
                v = float(r[vi].replace(",", ""))
            except Exception:
                continue
            unit = r[ui]
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit)
            kern = "ffa_bwd_kernel" if "ffa_bwd_kernel" in r[ki] else "ffa_fwd_kernel" if "ffa_fwd_kernel" in r[ki] else r[ki][:40]
            d = out.setdefault(kern, {})
            if mult is not None:
                d[r[mi]] = v * mult
            else:
                d[r[mi] + "_" + unit] = v
        for kern, d in out.items():
            d["traffic_bytes"] = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        out["_how"] = ("ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ffa_(fwd|bwd)_kernel "
                       "-c 2 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu   (seq 262144, 32 heads, per launch)")
        json.dump(out, open(os.path.join(OUT, "traffic_256k.json"), "w"), indent=1)

b = os.path.join(G, "bench_n1.json")
if os.path.exists(b) and os.path.getsize(b):
    with open(os.path.join(OUT, f"{tag}_bench_n1.json"), "w") as f:
        f.write(open(b).read())
print("profiles written:", sorted(os.listdir(OUT)))
