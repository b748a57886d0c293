"""B200-native CallableLoci hot path of DecodingUsTools' `coverage` command.

Layout (host-side mirror of the reference's `callable_loci` module; the device work lives in csrc/ behind the C ABI of
include/callable_loci_b200.h):

    options.py        CallableOptions, CalledState                  (reference: callable_loci/options.rs, types.rs)
    soa.py            ReadColumns: the packed read columns the device consumes
    callable_loci.py  CallableLociContext (ctypes over the C ABI), admission, CallableProfiler / ContigProfiler,
                      process_single_contig                          (reference: callable_loci/mod.rs, profilers/)
    sharding.py       region shards across GPUs, counter all-reduce, interval stitching
    report.py         natural contig order, build_coverage_export, summary.json, HTML page, SVG plot
    bam_stats.py      BAM sampler + sequencing-platform inference
    synth.py          synthetic workloads of the BASELINE configs (bench.py, tests)

Importing the package does not load the CUDA library; `callable_loci` does, and fails loudly if it has not been built.
"""
from .options import CallableOptions, CalledState
from .soa import ReadColumns

__all__ = ["CallableOptions", "CalledState", "ReadColumns"]
