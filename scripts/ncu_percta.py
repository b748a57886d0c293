"""Developer helper: warp-instructions per CTA by source line of clb_kernels.cuh."""
import csv, subprocess, io, sys
csv.field_size_limit(10**9)
rep = sys.argv[1]; ncta = int(sys.argv[2]); thr = float(sys.argv[3]) if len(sys.argv) > 3 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur = None; h = None; mine = []; tot = 0
def num(x):
    try: return int(x)
    except Exception: return 0
for r in rows:
    if r and r[0] == "File Path": cur = r[1]; continue
    if r and r[0] == "Line No": h = r; iI = h.index("Instructions Executed"); iS = h.index("# Samples"); continue
    if h and r and r[0].isdigit() and len(r) > iI:
        inst = num(r[iI]); samp = num(r[iS]); tot += inst
        if cur and cur.endswith("clb_kernels.cuh"): mine.append((int(r[0]), r[1], inst, samp))
print("total", tot, "per CTA", tot / ncta)
for ln, s, inst, samp in sorted(mine):
    if inst / ncta >= thr: print(f"{ln:4d} {inst / ncta:7.0f} {samp:5d} | {s.strip()[:130]}")
