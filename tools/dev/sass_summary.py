"""profiles/rNN_sass_summary.txt: per kernel of the shipped library, the SASS evidence of tcgen05 / TMEM / TMA use
(UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor load, UTMAREDG = TMA reduce-add, UTCBAR =
tcgen05.commit, SYNCS = mbarrier) and the resource usage (registers, spills, shared memory) from cuobjdump -res-usage."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "magiattention_b200/lib/libb200_ffa.so"
OPS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMAREDG", "UTMASTG", "UTCBAR", "SYNCS", "MUFU.EX2", "FFMA2", "RED.E", "ATOMG", "STS", "LDS"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    elif cur and "REG:" in line:
        usage[cur] = line.strip()
        cur = None
counts, name = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        counts[name] = collections.Counter()
        continue
    if name:
        for op in OPS:
            if re.search(r"\b" + re.escape(op), line):
                counts[name][op] += 1
dem = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
print(f"# {lib}: cuobjdump -sass / -res-usage summary (sm_100a)")
for n, c in counts.items():
    d = dem(n)
    short = re.sub(r"\(.*", "", d)
    print(f"\n== {short}")
    print("   " + "  ".join(f"{op}={c[op]}" for op in OPS if c[op]))
    if n in usage:
        print("   " + usage[n])
