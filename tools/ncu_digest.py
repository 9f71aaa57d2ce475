"""Digest of an .ncu-rep (run where ncu is installed): stall reasons, opcode mix per ray, hottest instructions.
usage: python tools/ncu_digest.py REPORT.ncu-rep [rays_per_launch] [kernel-substring]"""
import csv, io, subprocess, sys
from collections import Counter

rep = sys.argv[1]
rays = float(sys.argv[2]) if len(sys.argv) > 2 else 2048 * 100000
want = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
for r in rows[2:]:
    d = {h: r[i] for i, h in enumerate(hdr)}
    if want not in d["Kernel Name"]:
        continue
    print("=====", d["Kernel Name"][:90])
    keys = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    vals = sorted(((float(d[k].replace(",", "")), k) for k in keys if d[k]), reverse=True)
    print("  stalls/issue:", "  ".join("%s %.2f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for v, k in vals[:9]))
    for k in ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
              "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
              "smsp__inst_executed_op_shared_atom.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__warps_eligible.avg.per_cycle_active"]:
        if k in d:
            print("   %-70s %s" % (k, d[k]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
seen = set()
for b in blocks:
    if want not in b["name"] or b["name"] in seen:
        continue
    seen.add(b["name"])
    hdr, data = b["rows"][0], b["rows"][1:]
    isrc, isamp, iex, ith = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    idx = {h: hdr.index(h) for h in stall_cols}
    c, totex = Counter(), 0
    for r in data:
        if len(r) <= iex or not r[iex]:
            continue
        ex = int(r[iex]); totex += ex
        toks = r[isrc].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        c[op.split(".")[0]] += ex
    print("----", b["name"][:70], ": warp instructions", totex, " thread-instructions per ray ~ %.1f" % (totex * 32 / rays))
    print("   " + "  ".join("%s %.2f" % (op, v * 32 / rays) for op, v in c.most_common(30)))
    tot = sum(int(r[isamp] or 0) for r in data if len(r) > isamp)
    top = sorted(((int(r[isamp] or 0), i) for i, r in enumerate(data) if len(r) > isamp), reverse=True)[:int(sys.argv[4]) if len(sys.argv) > 4 else 24]
    print("   total samples", tot)
    for sm, i in sorted(top, key=lambda x: x[1]):
        r = data[i]
        st = sorted(((int(r[idx[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
        print("   %5d %6d %8dk thr%5s  %-58s %s" % (i, sm, int(r[iex] or 0) // 1000, r[ith][:5], r[isrc][:58], " ".join("%s:%d" % (h, v) for v, h in st if v > 0)))
