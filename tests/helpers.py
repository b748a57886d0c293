"""Shared helpers for the parity tests: run the same contigs through the oracle and the GPU path."""
from __future__ import annotations

import numpy as np

from decodingustools_b200 import report
from decodingustools_b200.callable_loci import (CallableLociContext, CallableProfiler, ContigProfiler,
                                                process_single_contig)
from oracle import oracle


def largest_non_chrm(contigs):
    return max([c[2] for c in contigs if c[0] != "chrM"] or [0])


def run_oracle(contigs, opt, debug=False):
    """contigs: list of (name, tid, length, ref_bytes_or_array, ReadColumns)."""
    o = oracle.OracleRun(opt, largest_non_chrm(contigs))
    for name, tid, length, ref, reads in contigs:
        o.process_contig(name, tid, length, ref, reads, debug=debug)
    return o


def run_gpu(contigs, opt, ctx=None, batch_reads=0):
    own = ctx is None
    ctx = ctx or CallableLociContext(opt)
    largest = largest_non_chrm(contigs)
    counter = CallableProfiler(None, largest)
    stats = {}
    results = []
    for name, tid, length, ref, reads in contigs:
        stats[tid] = ContigProfiler(name, length)
        results.append(process_single_contig(ctx, reads, ref, counter, stats, opt, tid, batch_reads=batch_reads))
    bed = counter.bed_bytes()
    export = report.build_coverage_export(stats, counter)
    if own:
        ctx.close()
    return bed, stats, counter, results, export


def assert_parity(contigs, opt, ctx=None, batch_reads=0):
    """BED bytes identical; all integer counters and bins bit-exact; floats equal (same f64 ops)."""
    o = run_oracle(contigs, opt)
    bed, stats, counter, results, export = run_gpu(contigs, opt, ctx, batch_reads)
    for (name, tid, length, _, _), oc in zip(contigs, o.contigs):
        s = stats[tid]
        assert counter.get_contig_counts(name).tolist() == oc.counts, (name, counter.get_contig_counts(name).tolist(), oc.counts)
        for k in ("n_covered_bases", "summed_coverage", "summed_baseq", "summed_mapq", "quality_bases", "n_reads"):
            assert getattr(s, k) == getattr(oc, k), (name, k, getattr(s, k), getattr(oc, k))
        assert (s.bins is None) == (oc.bins is None), name
        if oc.bins is not None:
            assert s.stride == oc.stride
            assert np.array_equal(s.bins, oc.bins), name
    assert bed == o.bed(), _first_diff(bed, o.bed())
    order, fl, sm = o.export()
    for c, f in zip(export["contigs"], fl):
        assert c["coverage_percent"] == f["coverage_percent"] and c["average_depth"] == f["average_depth"]
        assert c["quality_stats"] == {k: f[k] for k in ("average_mapq", "average_baseq", "q30_percentage")}
    assert export["summary"]["average_depth"] == sm["average_depth"]
    assert export["quality_metrics"] == {k: sm[k] for k in ("average_mapq", "average_baseq", "q30_percentage")}
    assert export["total_unique_reads"] == sm["total_unique_reads"]
    return o, results


def _first_diff(a: bytes, b: bytes) -> str:
    la, lb = a.splitlines(), b.splitlines()
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            return f"line {i}: gpu={x!r} oracle={y!r} (gpu lines {len(la)}, oracle lines {len(lb)})"
    return f"length differs: gpu lines {len(la)}, oracle lines {len(lb)}"
