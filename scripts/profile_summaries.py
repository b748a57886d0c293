"""Developer helper: turn the two ncu outputs of a round into the files under profiles/.

    python scripts/profile_summaries.py gpurun_out/r2_final_launches.csv gpurun_out/prof_r2_final.ncu-rep r02 "<command>" "<commit note>"

  <tag>_launches_raw.csv          the launch list as ncu wrote it (gpu__time_duration.sum per launch)
  <tag>_launch_list_summary.csv   per-kernel totals and shares, and the last HBM-resident step in launch order
  <tag>_k_pileup_fast_traffic.json   dram bytes / duration / grid / instructions of the --set full capture (bench.py reads roofline.traffic from it)
  <tag>_k_pileup_fast_resident_chr1.txt   raw metrics, stall shares, SASS regions and per-source-line instruction shares of that capture
"""
import csv, io, json, os, subprocess, sys
csv.field_size_limit(10**9)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launches, rep, tag, command, note = sys.argv[1:6]
out = os.path.join(ROOT, "profiles")

# ---- launch list
text = "".join(l for l in open(launches) if not l.startswith("=="))
open(os.path.join(out, f"{tag}_launches_raw.csv"), "w").write(text)
rows = list(csv.DictReader(io.StringIO(text)))
rows = [r for r in rows if r["Metric Name"] == "gpu__time_duration.sum"]
def us(r):
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    return v / 1e3 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1e3 if u in ("ms", "msecond") else v * 1e6
def short(n):
    n = n.replace("void ", "")
    return n.split("(")[0][:100]
tot, cnt = {}, {}
for r in rows:
    k = short(r["Kernel Name"]); tot[k] = tot.get(k, 0.0) + us(r); cnt[k] = cnt.get(k, 0) + 1
allt = sum(tot.values()); own = sum(v for k, v in tot.items() if k.startswith("clb::"))
with open(os.path.join(out, f"{tag}_launch_list_summary.csv"), "w") as f:
    f.write("kernel,launches,total_us,share_of_all_pct,share_of_own_kernels_pct\n")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        f.write(f"\"{k}\",{cnt[k]},{v:.1f},{100 * v / allt:.2f},{100 * v / own if k.startswith('clb::') else 0:.2f}\n")
    # the last resident step: from the last k_window_ranges launch on
    def grid(r): return int(r["Grid Size"].strip("()").split(",")[0])
    fast = [i for i, r in enumerate(rows) if "k_pileup_fast" in r["Kernel Name"]]
    full = max(grid(rows[i]) for i in fast)                      # a launch over ALL windows = a resident step (uploads launch windows in pieces)
    last = max(i for i in fast if grid(rows[i]) == full)
    idx = max(i for i in range(last) if "k_window_ranges" in rows[i]["Kernel Name"])
    end = min(i for i in range(last + 1, len(rows)) if "k_pack_stats" in rows[i]["Kernel Name"]) + 1    # k_pack_stats closes a step
    step = [r for r in rows[idx:end] if short(r["Kernel Name"]).startswith("clb::")]
    st = sum(us(r) for r in step)
    f.write("\n# one HBM-resident step (the last one of the run), in launch order\nkernel,us,share_of_step_pct,grid\n")
    for r in step:
        f.write(f"\"{short(r['Kernel Name'])}\",{us(r):.1f},{100 * us(r) / st:.2f},\"{r['Grid Size']}\"\n")

# ---- full capture
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r3 = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = r3[0], r3[1], r3[2]
g = lambda name: vals[hdr.index(name)]
def to_bytes(name):
    v = float(g(name).replace(",", "")); u = units[hdr.index(name)]
    return int(round(v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]))
def to_ns(name):
    v = float(g(name).replace(",", "")); u = units[hdr.index(name)]
    return int(round(v * {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}[u]))
traffic = {"dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
           "kernel": g("Kernel Name"), "grid": g("launch__grid_size"), "duration_ns": to_ns("gpu__time_duration.sum"),
           "warp_instructions": g("smsp__inst_executed.sum").split(".")[0],
           "workload": "chr1-size synthetic 30x 2x150bp, one HBM-resident launch over all windows (bench.py default workload)",
           "command": command, "commit": note}
json.dump(traffic, open(os.path.join(out, f"{tag}_k_pileup_fast_traffic.json"), "w"), indent=1)
a = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_fast.py"), rep, "1.0"], capture_output=True, text=True).stdout
b = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_lines.py"), rep, "50"], capture_output=True, text=True).stdout
with open(os.path.join(out, f"{tag}_k_pileup_fast_resident_chr1.txt"), "w") as f:
    f.write(f"# ncu --set full of ONE HBM-resident launch of {traffic['kernel']} on the chr1-size workload ({traffic['grid']} windows)\n")
    f.write(f"# command: {command}\n# {note}\n# summary made with scripts/profile_summaries.py (scripts/ncu_fast.py + scripts/ncu_lines.py)\n\n")
    f.write(a)
    f.write("\n# instructions and stall samples per CUDA source line (clb_fast.cuh unless the text says otherwise)\n")
    f.write("\n".join(l for l in b.splitlines() if "inst" in l and ("%" in l or "total" in l)))
    f.write("\n")
print(json.dumps(traffic, indent=1))
