"""Extract measured DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the trace kernels from an
``ncu --set full`` report and write ``profiles/trace_traffic.json``, which ``bench.py`` reports as ``roofline.traffic``.
usage: python tools/ncu_traffic.py <file.ncu-rep> <rays_per_launch> <source note>"""
import csv
import io
import json
import os
import subprocess
import sys

rep, rays, note = sys.argv[1], int(float(sys.argv[2])), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {"rays_per_launch": rays, "source": note, "kernels": {}}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"]
    key = "ab200_trace_fwd" if "trace_fwd" in name else "ab200_trace_bwd" if "trace_bwd" in name else None
    if key is None:
        continue
    total = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        total += float(d[m].replace(",", "")) * scale[units[hdr.index(m)]]
    out["kernels"][key] = {"dram_bytes_per_launch": total, "duration_ms_under_ncu": float(d["gpu__time_duration.sum"].replace(",", "")) *
                           {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units[hdr.index("gpu__time_duration.sum")]],
                           "warp_instructions_per_launch": float(d["smsp__inst_executed.sum"].replace(",", ""))}
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "trace_traffic.json")
with open(dst, "w") as fh:
    json.dump(out, fh, indent=1)
print(json.dumps(out, indent=1))
