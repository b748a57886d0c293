"""Developer helper: per-contig fixed cost of the C-ABI path (hg38 has 3366 contigs, most of them tiny)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from decodingustools_b200 import synth
from decodingustools_b200.callable_loci import CallableLociContext
from decodingustools_b200.options import CallableOptions

ctx = CallableLociContext(CallableOptions())
for length, depth in ((2_000, 30.0), (40_000, 30.0), (200_000, 30.0), (40_000, 0.0)):
    c = synth.synth_short("chrUn", length, seed=3, depth=depth) if depth else None
    name, ref, reads = ("chrUn", c.ref, c.reads) if c else ("chrUn", b"N" * length, None)
    n = 300
    for it in range(n + 20):
        if it == 20:
            t0 = time.perf_counter()
        ctx.begin_contig(0, name, length, ref, 248_956_422, max_ref_span=reads.max_ref_span() if reads is not None else 0)
        if reads is not None and reads.n:
            ctx.push_reads(reads)
        ctx.finish_contig(copy_intervals=False)
    dt = (time.perf_counter() - t0) / n
    print(f"contig {length} bp depth {depth}: {dt * 1e6:.0f} us per contig (begin + push + finish)")
