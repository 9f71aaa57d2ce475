"""Summarise an .ncu-rep (read on the CPU box): key raw metrics per kernel + SASS opcode mix per ray of the hot kernel.
usage: python tools/ncu_summary.py <file.ncu-rep> <rays_per_launch> [kernel-regex ...]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, rays = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.avg",
        "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
        "smsp__average_warp_latency_issue_stalled_no_instruction.ratio", "smsp__average_warp_latency_issue_stalled_branch_resolving.ratio",
        "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=" * 100)
    print(d.get("Kernel Name"))
    for k in want:
        if k in d:
            print(f"  {k:85s} {d[k]:>16s} {units[hdr.index(k)]}")
    if "smsp__thread_inst_executed.sum" in d:
        print(f"  thread instructions per ray: {float(d['smsp__thread_inst_executed.sum'].replace(',', '')) / rays:.1f}")
for pat in sys.argv[3:]:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
    rws = list(csv.reader(io.StringIO(src)))
    # several launches are concatenated; take the first block
    blocks, cur = [], None
    for r in rws:
        if r and r[0] == "Kernel Name":
            cur = []
            blocks.append((r[1], cur))
        elif cur is not None:
            cur.append(r)
    name, blk = blocks[0]
    h = blk[0]
    ia, ie = h.index("Source"), h.index("Thread Instructions Executed")
    ops = collections.Counter()
    tot = 0
    for r in blk[1:]:
        if len(r) <= ie or not r[ie].isdigit():
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ia])
        op = m.group(2) if m else "?"
        ops[op] += int(r[ie])
        tot += int(r[ie])
    print("=" * 100)
    print(f"SASS opcode mix of {name[:80]} (first profiled launch): {tot / rays:.1f} thread instructions per ray")
    for op, c in ops.most_common(28):
        print(f"  {op:10s} {c / rays:7.2f} per ray  {100 * c / tot:5.1f} %")
