"""Developer helper: extra GPU-vs-oracle parity runs beyond the test suite (random micro-contigs with random options, and
mid-size synthetic contigs with other read lengths and depths near the fast kernel's 254-read limit).
    python scripts/stress_parity.py [first_seed] [n_seeds]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from decodingustools_b200 import synth
from decodingustools_b200.options import CallableOptions
from tests.helpers import assert_parity
from decodingustools_b200.callable_loci import CallableLociContext
from tests.test_oracle_vs_naive import random_reads
bad = 0
seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
nseeds = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for seed in range(seed0, seed0 + nseeds):
    rng = np.random.default_rng(seed)
    length = int(rng.integers(1, 9000))
    ref = bytes(rng.choice(list(b"ACGTNnR"), size=length, p=[.2, .2, .2, .2, .1, .05, .05]).tolist())
    reads = random_reads(rng, length, int(rng.integers(0, 600)), max_len=int(rng.choice([30, 150, 400])))
    opt = CallableOptions(min_depth=int(rng.integers(0, 40)), max_depth=int(rng.choice([0, 2, 30, 100, 254, 500])),
                          min_mapping_quality=int(rng.choice([0, 10, 30])), min_base_quality=int(rng.choice([0, 20, 200])),
                          min_depth_for_low_mapq=int(rng.integers(0, 30)), max_low_mapq=int(rng.choice([0, 1, 9])),
                          max_low_mapq_fraction=float(rng.choice([-0.5, 0.0, 0.1, 0.25, 0.5, 1.0])))
    try:
        ctx = CallableLociContext(opt)
        try:
            assert_parity([("chrT", int(rng.integers(0, 3)), length, ref, reads)], opt, ctx)
        finally:
            ctx.close()
    except AssertionError as e:
        bad += 1; print("FAIL seed", seed, str(e)[:200])
# mid-size synthetic contigs with different read lengths / depths (fast kernel sub-batch sizing, X list, bail paths)
ctx = CallableLociContext(CallableOptions())
for i, (rl, depth) in enumerate([(36, 60.0), (75, 30.0), (100, 15.0), (151, 45.0), (250, 30.0), (150, 230.0), (150, 300.0)]):
    c = synth.synth_short("chr7", 300_000, seed=900 + i, depth=depth, read_len=rl)
    try:
        o, res = assert_parity([(c.name, 0, c.length, c.ref, c.reads)], CallableOptions(), ctx)
        print("ok", rl, depth, "general windows", res[0].general_windows)
    except AssertionError as e:
        bad += 1; print("FAIL synth", rl, depth, str(e)[:200])
print("failures:", bad)
