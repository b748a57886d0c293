"""Developer helper: one contig as region shards on ONE GPU against the whole-contig run (what bench.py's strong leg does over N GPUs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from decodingustools_b200 import synth, sharding, _lib
from decodingustools_b200.callable_loci import CallableLociContext, stitch_intervals
from decodingustools_b200.options import CallableOptions
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2
world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
opt = CallableOptions()
c = synth.synth_short("chr1", int(synth.HG38["chr1"] * scale), synth.SEED0 + 1)
reads = c.reads; span = reads.max_ref_span()
ctx = CallableLociContext(opt)
ctx.begin_contig(0, "chr1", c.length, c.ref, c.length, max_ref_span=span); ctx.push_reads(reads); whole = ctx.finish_contig()
W = int(_lib.lib().clb_window_positions())
parts = []
for r in range(world):
    sh = sharding.plan_regions([c.length], world, W)[r][0]
    lo, hi = sharding.reads_for_region(reads, sh.start, sh.end, span)
    ctx2 = CallableLociContext(opt)
    ctx2.begin_contig(0, "chr1", c.length, c.ref, c.length, region=(sh.start, sh.end), max_ref_span=span)
    ctx2.push_reads(reads.slice(lo, hi)); p = ctx2.finish_contig(); parts.append(p)
    print("shard", r, sh, "reads", hi - lo, "general windows", p.general_windows, "counts", p.state_counts.tolist())
    ctx2.close()
tot = sum(p.state_counts for p in parts)
print("whole ", whole.state_counts.tolist()); print("shards", tot.tolist())
iv = stitch_intervals([p.intervals for p in parts])
print("intervals", iv.shape, whole.intervals.shape)
n = min(iv.shape[0], whole.intervals.shape[0])
d = np.flatnonzero((iv["start"][:n] != whole.intervals["start"][:n]) | (iv["state"][:n] != whole.intervals["state"][:n]))
print("first diff", d[:3], iv[d[:3]] if d.size else None, whole.intervals[d[:3]] if d.size else None)
