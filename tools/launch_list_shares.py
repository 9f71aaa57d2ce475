"""Per-kernel share of a step from an ncu launch list (``--metrics gpu__time_duration.sum --csv --log-file``).
usage: python tools/launch_list_shares.py profiles/r02_ncu_launch_list.csv"""
import collections
import csv
import io
import sys

lines = open(sys.argv[1]).read().splitlines()
start = next(i for i, ln in enumerate(lines) if ln.startswith('"ID"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
t = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r["Kernel Name"]
    short = name.split("(")[0].split("::")[-1][:60] if "ab200" in name or "trace_" in name or "nurbs_" in name else "torch: " + name.split("<")[0][-40:]
    t[short][0] += 1
    t[short][1] += float(r["Metric Value"]) / 1e6
tot = sum(v[1] for v in t.values())
print(f"{len(rows)} launches, {tot:.3f} ms of kernel time under ncu (cold caches, serialised)")
for k, v in sorted(t.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{v[1]:9.3f} ms  {v[0]:4d}x  {100 * v[1] / tot:5.1f} %  {k}")
