"""Parity workloads under a -DCLB_BOUNDS_CHECK build of the library (CLB_LIB must point at it): every shared-memory atomic /
vector load and every streamed quality load of the kernels is range-checked on the device.  Prints one JSON line with
the number of violations (must be 0).  Stand-in for compute-sanitizer memcheck, which the GPU pool does not allow."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from decodingustools_b200 import _lib, synth
from decodingustools_b200.callable_loci import CallableLociContext
from decodingustools_b200.options import CallableOptions
from decodingustools_b200.soa import ReadColumns
from tests.helpers import assert_parity
from tests.test_oracle_vs_naive import random_reads

L = _lib.lib()
W = int(L.clb_window_positions())
cases = 0
for seed in range(12):                                       # random micro contigs, odd options
    rng = np.random.default_rng(1000 + seed)
    length = int(rng.integers(1, 60))
    ref = bytes(rng.choice(list(b"ACGTNnR"), size=length, p=[.2, .2, .2, .2, .1, .05, .05]).tolist())
    opt = CallableOptions(min_depth=int(rng.integers(0, 5)), max_depth=int(rng.choice([0, 2, 3, 5, 500])), min_base_quality=int(rng.choice([0, 20, 200])))
    assert_parity([("chrT", 0, length, ref, random_reads(rng, length, int(rng.integers(0, 80))))], opt); cases += 1
for length in (1, W - 1, W, W + 1, 2 * W + 1, 3 * W - 1):    # reads on window seams
    rng = np.random.default_rng(length)
    ref = bytes(rng.choice(list(b"ACGTN"), size=length, p=[.24, .24, .24, .24, .04]).tolist())
    recs = []
    for seam in (0, W, 2 * W):
        for d in (-151, -150, -2, -1, 0, 1):
            p = seam + d
            if 0 <= p and p + 150 <= length:
                recs += [(p, 0, 60, "150M", rng.choice([2, 37], size=150).tolist(), f"s{seam}_{d}_{k}") for k in range(4)]
                recs.append((p, 0, 60, "20S50M10D50M30S", rng.choice([2, 37], size=150).tolist(), f"c{seam}_{d}"))
    recs.sort(key=lambda r: r[0])
    assert_parity([("chrE", 0, length, ref, ReadColumns.from_records(recs))], CallableOptions()); cases += 1
c = synth.synth_short("chr22", 400_000, seed=11); assert_parity([(c.name, 0, c.length, c.ref, c.reads)], CallableOptions(), batch_reads=7777); cases += 1
c = synth.synth_long("chr1", 300_000, seed=13); assert_parity([(c.name, 0, c.length, c.ref, c.reads)], CallableOptions()); cases += 1
c = synth.synth_short("chrY", 30_000, seed=14, depth=2000.0)
assert_parity([(c.name, 0, c.length, c.ref, c.reads)], CallableOptions()); assert_parity([(c.name, 0, c.length, c.ref, c.reads)], CallableOptions(max_depth=4000)); cases += 2
recs = [(100 + (i % 3), 0, 60, "100M", 30, f"d{i}") for i in range(70_000)]; recs.sort(key=lambda r: r[0])      # > 65535 reads in one window
assert_parity([("chrD", 0, 5000, bytes(b"ACGT" * 1250), ReadColumns.from_records(recs))], CallableOptions(max_depth=100000)); cases += 1
ctx = CallableLociContext(CallableOptions())
v, checked = C.c_uint32(0), C.c_int(0)
L.clb_debug_bounds.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
assert L.clb_debug_bounds(ctx._h, C.byref(v), C.byref(checked)) == 0
ctx.close()
print(json.dumps({"library": os.environ.get("CLB_LIB", "default"), "bounds_checked_build": bool(checked.value), "parity_cases": cases,
                  "out_of_bounds_accesses_caught": int(v.value)}))
