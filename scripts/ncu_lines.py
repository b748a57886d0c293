"""Developer helper: summarise an .ncu-rep (kernel totals + instructions per CUDA source line)."""
import csv, subprocess, sys, io
csv.field_size_limit(10**9)
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for i, h in enumerate(hdr):
    if h in want or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        print(f"{h:90s} {units[i]:12s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
lines = []
for r in rows:
    if r and r[0] == "Line No":
        h = r; iI = h.index("Instructions Executed"); iS = h.index("# Samples"); continue
    if h and r and r[0].isdigit():
        try:
            lines.append((int(r[0]), r[1], int(r[iI]) if r[iI] not in ("-", "") else 0, int(r[iS]) if r[iS] not in ("-", "") else 0))
        except Exception:
            pass
tot = sum(l[2] for l in lines) or 1; ts = sum(l[3] for l in lines) or 1
print("total line-attributed inst", tot, "samples", ts)
for ln, s, inst, samp in sorted(lines, key=lambda x: -x[2])[:topn]:
    print(f"{ln:4d} {100 * inst / tot:5.1f}% inst {100 * samp / ts:5.1f}% samp | {s.strip()[:120]}")
