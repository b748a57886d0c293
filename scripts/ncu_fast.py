"""Developer helper: key metrics, stall reasons and hot SASS lines of an .ncu-rep capture of one kernel."""
import csv, subprocess, sys, io
csv.field_size_limit(10**9)
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:90s} {units[i]:12s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; iS = h.index("Source"); iN = h.index("# Samples"); iI = h.index("Instructions Executed")
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
st = {k: h.index(k) for k in stalls}
data = [(r[iS].strip(), int(r[iN]), int(r[iI]), {k: int(r[v]) for k, v in st.items()}) for r in rows[2:] if len(r) >= len(h)]
tot = sum(d[1] for d in data) or 1; toti = sum(d[2] for d in data) or 1
agg = {}
for d in data:
    for k, v in d[3].items(): agg[k] = agg.get(k, 0) + v
print("stall shares %:", {k[6:]: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v * 100 / tot >= 0.5})
marks = [i for i, d in enumerate(data) if any(t in d[0] for t in ("UBLKCP", "SYNCS", "BAR.SYNC"))]
print("markers:", [(i, data[i][0][:28], data[i][2]) for i in marks])
prev = 0
for m in marks + [len(data)]:
    if m > prev:
        print(f"  [{prev:4d},{m:4d}) inst {100 * sum(d[2] for d in data[prev:m]) / toti:5.1f}%  samples {100 * sum(d[1] for d in data[prev:m]) / tot:5.1f}%")
    prev = m
for idx, (s, n, i, sd) in enumerate(data):
    if n * 100 / tot >= thr:
        top = sorted(sd.items(), key=lambda x: -x[1])[:2]
        print(f"{idx:5d} {100 * n / tot:5.1f}% {i:9d} {s[:64]:64s} {[(k[6:], v) for k, v in top]}")
