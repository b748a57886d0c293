"""The golden cases of tests/golden/ and how their expected outputs are derived (shared by scripts/make_golden_bundle.py,
which writes the bundle, and tests/test_golden.py, which checks the oracle and the CUDA path against it).
TEST INFRASTRUCTURE: uses the CPU oracle."""
from __future__ import annotations

import hashlib

import numpy as np

from decodingustools_b200 import bam_stats, report, synth
from decodingustools_b200.callable_loci import ContigProfiler
from decodingustools_b200.options import CallableOptions
from decodingustools_b200.soa import ReadColumns
from oracle import oracle

REF10 = b"NNACGTACGT"
FLAG_NAMES = [("min_depth", "--min-depth", 4), ("max_depth", "--max-depth", 500), ("min_mapping_quality", "--min-mapping-quality", 10),
              ("min_base_quality", "--min-base-quality", 20), ("min_depth_for_low_mapq", "--min-depth-for-low-mapq", 10),
              ("max_low_mapq", "--max-low-mapq", 1), ("max_low_mapq_fraction", "--max-low-mapq-fraction", 0.1)]


def qname(contig: str, name_id: int) -> str:
    """Illumina-style read names (mates share one), so that the BamStats sampler infers a platform."""
    return f"A00123:7:HFLOWCELLX:1:1101:{len(contig)}:{name_id}"


def cases():
    """(case name, [(contig name, length, ref bytes, ReadColumns)], CallableOptions) -- SURVEY.md section 4 KA1-KA7, a
    depth-cap pile, and miniatures of BASELINE configs 1, 4 and 5."""
    R = ReadColumns.from_records
    four = [(2, 0, 60, "5M", 30, f"r{i}") for i in range(4)]
    yield "ka1_no_reads", [("c1", 10, REF10, ReadColumns.empty())], CallableOptions()
    yield "ka2_four_reads", [("c1", 10, REF10, R(four))], CallableOptions()
    yield "ka3_deletion", [("c1", 10, REF10, R([(2, 0, 60, "2M1D2M", 30, f"r{i}") for i in range(4)]))], CallableOptions()
    yield "ka4_low_mapq_two_of_ten", [("c1", 10, REF10, R([(2, 0, 0 if i < 2 else 60, "5M", 30, f"r{i}") for i in range(10)]))], CallableOptions()
    yield "ka4_low_mapq_one_of_ten", [("c1", 10, REF10, R([(2, 0, 0 if i < 1 else 60, "5M", 30, f"r{i}") for i in range(10)]))], CallableOptions()
    yield "ka5_contig_boundary_quirk", [("a", 3, b"ACG", ReadColumns.empty()), ("b", 2, b"AC", ReadColumns.empty())], CallableOptions()
    yield ("ka6_depth_cap_3", [("c", 6, b"ACGTAC", R([(0, 0, 60, "5M", 30, f"a{i}") for i in range(5)] + [(1, 0, 60, "5M", 30, f"b{i}") for i in range(2)]))],
           CallableOptions(max_depth=3, min_depth=1))
    yield "ka7_ref_n_lowercase_iupac", [("c1", 10, b"nnRCGTACGT", R(four))], CallableOptions()
    # more than 500 records at and around one position: htslib's admission (first-at-position bypass, overshoot)
    rng = np.random.default_rng(600)
    pile = []
    for p, n in ((100, 5), (120, 620), (121, 40), (122, 3), (180, 700), (181, 2), (400, 10)):
        for i in range(n):
            pile.append((p, int(rng.choice([0, 0x400, 0x100, 4], p=[.9, .05, .03, .02])), int(rng.choice([60, 0, 5], p=[.9, .05, .05])),
                         str(rng.choice(["100M", "40M5D60M", "10S90M", "50M3I47M"])), rng.choice([2, 23, 37], size=100, p=[.05, .1, .85]).tolist(), f"p{p}_{i}"))
    ref = bytes(np.random.default_rng(601).choice(list(b"ACGT"), size=700).tolist())
    yield "depth_cap_pile_default_500", [("chrP", 700, ref, R(pile))], CallableOptions()
    # BASELINE configs[0]/[1] miniature: two 30x contigs + chrM
    cs = [synth.synth_short("chr21", 30_000, seed=9001), synth.synth_short("chr22", 12_000, seed=9002), synth.synth_short("chrM", 16_569, seed=9003, depth=60.0)]
    yield "mini_config1_30x", [(c.name, c.length, c.ref.tobytes(), c.reads) for c in cs], CallableOptions()
    # BASELINE configs[3] miniature: a deep pile with the default cap, and with the cap lifted
    d = synth.synth_short("chrY", 1_500, seed=9004, depth=1500.0)
    yield "mini_config4_deep_cap500", [(d.name, d.length, d.ref.tobytes(), d.reads)], CallableOptions()
    yield "mini_config4_deep_maxdepth4000", [(d.name, d.length, d.ref.tobytes(), d.reads)], CallableOptions(max_depth=4000)
    # BASELINE configs[4] miniature: long indel-heavy reads
    lg = synth.synth_long("chr1", 40_000, seed=9005, depth=20.0)
    yield "mini_config5_long_reads", [(lg.name, lg.length, lg.ref.tobytes(), lg.reads)], CallableOptions()


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def expected_outputs(contigs, opt, header_text, names_by_tid):
    """BED bytes and summary.json text from the oracle (+ the BamStats sampler mirror for the four header fields)."""
    largest = max([l for n, l, _, _ in contigs if n != "chrM"] or [0])
    o = oracle.OracleRun(opt, largest)
    stats, counts = {}, {}
    for tid, (name, length, ref, reads) in enumerate(contigs):
        oc = o.process_contig(name, tid, length, ref, reads)
        s = ContigProfiler(name, length, oc.n_covered_bases, oc.summed_coverage, oc.summed_baseq, oc.summed_mapq, oc.quality_bases, oc.n_reads)
        stats[tid] = s; counts[name] = np.array(oc.counts, np.uint64)

    class _Counter:
        def get_contig_counts(self, n):
            return counts[n]
    bs = bam_stats.BamStats()
    bs.set_header(header_text)
    done = False
    for tid, (name, length, ref, reads) in enumerate(contigs):
        for i in range(reads.n):
            if not bs.add_record(names_by_tid[tid][i], int(reads.flag[i]), int(reads.qual_off[i + 1] - reads.qual_off[i])):
                done = True
                break
        if done:
            break
    export = report.build_coverage_export(stats, _Counter(), bs.as_summary_fields())
    plots = [f"{oc.name}_coverage.svg" for oc in o.contigs if oc.bins is not None]
    return o.bed(), report.render_summary_json(export, "callable_regions.bed", "summary.html", plots)


