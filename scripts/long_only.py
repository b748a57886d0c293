"""Developer helper: one resident run of the long-read config (target for an ncu capture)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from decodingustools_b200 import synth
from decodingustools_b200.options import CallableOptions
from config_times import run

run("5: long reads 15kb indel-heavy, 25 Mbp", synth.synth_long("chr1", 25_000_000, 5), CallableOptions())
