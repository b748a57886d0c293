"""Developer helper: where does a window's wall time go? (clock64 stamps per CTA via clb_debug_timing)

Needs a library built with the stamps: (cd decodingustools_b200/csrc && nvcc <Makefile flags> -DCLB_PHASE_TIMING -shared
-o ../../variants/lib_timing.so callable_loci_b200.cu clb_host.cpp -ldl), then CLB_LIB=variants/lib_timing.so python scripts/phase_timing.py."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from decodingustools_b200 import synth, _lib
from decodingustools_b200.callable_loci import CallableLociContext, admit_reads, compact_reads
from decodingustools_b200.options import CallableOptions
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
mode = sys.argv[2] if len(sys.argv) > 2 else "short"
c = synth.synth_short("chr1", int(synth.HG38["chr1"] * scale), 1) if mode == "short" else synth.synth_long("chr1", int(synth.HG38["chr1"] * scale), 1)
reads = compact_reads(c.reads, admit_reads(c.reads, 500, 0))
ctx = CallableLociContext(CallableOptions())
ctx.begin_contig(0, "chr1", c.length, c.ref, c.length, max_ref_span=reads.max_ref_span())
ctx.push_reads(reads); ctx.finish_contig(copy_intervals=False)
L = _lib.lib()
nmax = 200000
buf = np.zeros((nmax, 8), np.int64); n = C.c_uint32(0)
L.clb_debug_timing.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
rc = L.clb_debug_timing(ctx._h, buf.ctypes.data_as(C.c_void_p), nmax, C.byref(n))
assert rc == 0
t = buf[: n.value]
busy = t[t[:, 7] > (300 if mode == 'short' else 20)]          # windows with a normal read load
names = (["setup (tables, zeroing, first columns)", "warp 0: first CIGAR walk", "warp 0: wait for its bulk copy", "warp 0: first stream",
          "rest of the loop + barriers + extras", "phase C"] if not os.environ.get("CLB_FORCE_GENERAL") else
         ["setup", "phaseA(round0)", "phaseB+rest rounds", "C: scan+classify+stats", "C: boundaries+records", "C: bins"])
d = np.diff(busy[:, :7], axis=1).astype(np.float64)
tot = d.sum(axis=1)
print(f"windows {n.value}, busy {busy.shape[0]}, mean candidates {busy[:,7].mean():.0f}, mean cycles/window {tot.mean():.0f}")
for i, nm in enumerate(names):
    print(f"  {nm:44s} {d[:, i].mean():9.0f} cyc  {100 * d[:, i].mean() / tot.mean():5.1f}%")
