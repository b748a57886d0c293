"""Developer helper: instructions per source file and per marked region of clb_kernels.cuh."""
import csv, subprocess, sys, io, re
csv.field_size_limit(10**9)
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur = None; h = None; per_file = {}; mine = []
for r in rows:
    if r and r[0] == "File Path": cur = r[1]; continue
    if r and r[0] == "Line No": h = r; iI = h.index("Instructions Executed"); continue
    if h and r and r[0].isdigit():
        inst = int(r[iI]) if r[iI] not in ("-", "") else 0
        per_file[cur] = per_file.get(cur, 0) + inst
        if cur and cur.endswith("clb_kernels.cuh"): mine.append((int(r[0]), r[1], inst))
tot = sum(per_file.values())
for f, v in sorted(per_file.items(), key=lambda x: -x[1]): print(f"{100*v/tot:5.1f}%  {f}")
# regions by function markers in my file
text = open("decodingustools_b200/csrc/clb_kernels.cuh").read().splitlines()
marks = []
for i, l in enumerate(text, 1):
    for name, pat in [("bytes_lt", "uint32_t bytes_lt("), ("emit_m", "bool emit_m("), ("emit_read", "void emit_read("), ("process_chunk", "void process_chunk("),
                      ("process_slots", "void process_slots("), ("run_segments", "void run_segments("), ("red_helpers", "uint32_t smem_addr("), ("kernel_setup", "k_pileup_classify(const KParams P)"), ("phaseA", "phase A/B: reads -> counters"),
                      ("complex_path", "long CIGARs: the whole warp"), ("phaseC_scan", "phase C: scan, classify"), ("classify", "uint32_t st[PPT];"),
                      ("boundaries", "// run boundaries"), ("bins", "// bins: positions of"), ("stats", "// per-CTA reduction of the additive"), ("helpers", "// Small helper kernels")]:
        if pat in l: marks.append((i, name))
marks.sort()
def region(ln):
    name = "top"
    for i, n in marks:
        if ln >= i: name = n
    return name
agg = {}
for ln, s, inst in mine: agg[region(ln)] = agg.get(region(ln), 0) + inst
for k, v in sorted(agg.items(), key=lambda x: -x[1]): print(f"   {100*v/tot:5.1f}%  {k}")
