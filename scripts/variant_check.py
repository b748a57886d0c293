"""Developer helper: quick parity check of the library selected by CLB_LIB against the oracle (small synthetic contig)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from decodingustools_b200 import synth
from decodingustools_b200.options import CallableOptions
from tests.helpers import assert_parity
c = synth.synth_short("chr22", 400_000, seed=5)
l = synth.synth_long("chr1", 200_000, seed=6)
assert_parity([(c.name, 0, c.length, c.ref, c.reads), (l.name, 1, l.length, l.ref, l.reads)], CallableOptions())
print("parity ok", os.path.basename(os.environ.get("CLB_LIB", "default")))
