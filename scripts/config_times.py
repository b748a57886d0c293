"""Developer helper: HBM-resident kernel time of the other BASELINE configs at reduced scale (long reads, 2000x)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from decodingustools_b200 import synth
from decodingustools_b200.callable_loci import CallableLociContext, admit_reads, compact_reads
from decodingustools_b200.options import CallableOptions

def run(tag, c, opt):
    keep = admit_reads(c.reads, opt.pileup_max_depth, 0)
    reads = compact_reads(c.reads, keep)
    ctx = CallableLociContext(opt)
    ctx.begin_contig(0, c.name, c.length, c.ref, c.length, max_ref_span=reads.max_ref_span())
    ctx.push_reads(reads)
    r = ctx.finish_contig(copy_intervals=False)
    best = min(ctx.rerun_resident(fetch=True)[1].pileup_ms for _ in range(4))
    byts = reads.nbytes_device() + c.length // 8
    print(json.dumps({"config": tag, "reads": reads.n, "offered": c.reads.n, "cigar_ops": reads.n_cigar, "cells": r.summed_coverage, "pileup_ms": round(best, 3),
                      "Gcells_s": round(r.summed_coverage / best / 1e6, 1), "GBps": round(byts / best / 1e6, 1), "frac": round(byts / best / 1e6 / 6537, 4)}))
    ctx.close()

if __name__ == "__main__":
  run("5: long reads 15kb indel-heavy, 25 Mbp", synth.synth_long("chr1", 25_000_000, 5), CallableOptions())
  run("4: 2000x chrY-size/20 (cap 500)", synth.synth_short("chrY", 2_800_000, 4, depth=2000.0), CallableOptions())
  run("4b: 2000x, --max-depth 4000 (no cap)", synth.synth_short("chrY", 1_000_000, 4, depth=2000.0), CallableOptions(max_depth=4000))
  run("1: chr22-size 30x", synth.synth_short("chr22", synth.HG38["chr22"], 1), CallableOptions())
