"""Developer helper: HBM-resident kernel time of the current library (CLB_LIB selects a variant build)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from decodingustools_b200 import synth
from decodingustools_b200.callable_loci import CallableLociContext, admit_reads, compact_reads
from decodingustools_b200.options import CallableOptions

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mode = sys.argv[3] if len(sys.argv) > 3 else "short"
opt = CallableOptions()
L = int(synth.HG38["chr1"] * scale)
c = synth.synth_short("chr1", L, 1, read_len=int(os.environ.get("CLB_READ_LEN", "150"))) if mode == "short" else synth.synth_long("chr1", L, 1)
reads = compact_reads(c.reads, admit_reads(c.reads, 500, 0))
ctx = CallableLociContext(opt)
ctx.begin_contig(0, "chr1", c.length, c.ref, c.length, max_ref_span=reads.max_ref_span())
ctx.push_reads(reads)
try:
    r = ctx.finish_contig(copy_intervals=False)
except Exception as e:
    if not os.environ.get('CLB_TOLERATE'): raise
    class R: summed_coverage = int(reads.ref_len().sum())
    r = R()
ms = []
for i in range(steps + 2):
    try:
        t, res = ctx.rerun_resident(fetch=True)
        pm = res.pileup_ms
    except Exception:
        if not os.environ.get('CLB_TOLERATE'): raise
        t, _ = ctx.rerun_resident(fetch=False); pm = t
    if i >= 2: ms.append((t, pm, getattr(res, "fast_ms", 0.0), getattr(res, "general_windows", -1)))
best = min(m[1] for m in ms)
byts = reads.nbytes_device() + c.length // 8
print(json.dumps({"lib": os.path.basename(os.environ.get("CLB_LIB", "default")), "pileup_ms": best, "all_ms": min(m[0] for m in ms),
                  "GBps": byts / best / 1e6, "frac": byts / best / 1e6 / 6537, "fast_ms": min(m[2] for m in ms), "general_windows": ms[-1][3], "Gcells_s": r.summed_coverage / best / 1e6}))
