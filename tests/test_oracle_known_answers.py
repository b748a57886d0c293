"""Oracle pins: the hand-derived known-answer vectors KA1-KA7 of SURVEY.md section 4.

The reference has no tests or fixtures for this path (parity unpinned); these vectors were derived
by reading /root/reference/src/callable_loci/{mod.rs,profilers/*.rs} and are checked against BOTH
the C oracle (iterator mechanics) and the naive Python model (closed-form rules).
"""
import numpy as np
import pytest

from decodingustools_b200.options import CallableOptions
from decodingustools_b200.soa import ReadColumns
from oracle import naive_model, oracle

REF10 = b"NNACGTACGT"


def run_both(contigs, opt, largest=None):
    """contigs: list of (name, tid, length, ref_bytes, ReadColumns). Returns (OracleRun, NaiveRun)."""
    if largest is None:
        largest = max([c[2] for c in contigs if c[0] != "chrM"] or [0])
    o = oracle.OracleRun(opt, largest)
    n = naive_model.NaiveRun(opt, largest)
    for name, tid, length, ref, reads in contigs:
        o.process_contig(name, tid, length, ref, reads, debug=True)
        n.process_contig(name, length, ref, reads, tid=tid)
    assert o.bed() == n.bed()
    for oc, nc in zip(o.contigs, n.results):
        assert oc.counts == nc["counts"]
        assert list(oc.raw) == nc["raw"] and list(oc.qc) == nc["qc"] and list(oc.low) == nc["low"]
        for k in ("n_covered_bases", "summed_coverage", "summed_baseq", "summed_mapq", "quality_bases", "n_reads"):
            assert getattr(oc, k) == nc[k], k
        assert (oc.bins is None) == (nc["bins"] is None)
        if oc.bins is not None:
            assert oc.bins.tolist() == nc["bins"]
    return o, n


def test_ka1_no_reads():
    o, _ = run_both([("c1", 0, 10, REF10, ReadColumns.empty())], CallableOptions())
    assert o.bed() == b"c1\t0\t2\tREF_N\nc1\t2\t10\tNO_COVERAGE\n"
    c = o.contigs[0]
    assert c.counts == [2, 0, 8, 0, 0, 0]
    assert c.n_covered_bases == 0
    _, fl, sm = o.export()
    assert fl[0]["average_depth"] == 0.0 and sm["average_depth"] == 0.0


def test_ka2_four_reads():
    reads = ReadColumns.from_records([(2, 0, 60, "5M", 30, f"r{i}") for i in range(4)])
    o, _ = run_both([("c1", 0, 10, REF10, reads)], CallableOptions())
    assert o.bed() == b"c1\t0\t2\tREF_N\nc1\t2\t7\tCALLABLE\nc1\t7\t10\tNO_COVERAGE\n"
    c = o.contigs[0]
    assert (c.summed_coverage, c.n_covered_bases) == (20, 5)
    assert (c.quality_bases, c.summed_baseq, c.summed_mapq, c.n_reads) == (20, 600, 1200, 4)
    _, fl, _ = o.export()
    assert fl[0] == dict(coverage_percent=50.0, average_depth=4.0, average_mapq=60.0, average_baseq=30.0,
                         q30_percentage=100.0)


def test_ka3_deletion():
    reads = ReadColumns.from_records([(2, 0, 60, "2M1D2M", 30, f"r{i}") for i in range(4)])
    o, _ = run_both([("c1", 0, 10, REF10, reads)], CallableOptions())
    assert o.bed() == (b"c1\t0\t2\tREF_N\nc1\t2\t4\tCALLABLE\nc1\t4\t5\tLOW_COVERAGE\n"
                       b"c1\t5\t7\tCALLABLE\nc1\t7\t10\tNO_COVERAGE\n")
    c = o.contigs[0]
    assert list(c.raw[2:7]) == [4, 4, 4, 4, 4] and list(c.qc[2:7]) == [4, 4, 0, 4, 4]
    assert (c.summed_mapq, c.quality_bases) == (1200, 16)
    _, fl, _ = o.export()
    assert fl[0]["average_mapq"] == 75.0          # quirk Q5


def test_ka4_low_mapq_fraction():
    two_low = ReadColumns.from_records([(2, 0, 0 if i < 2 else 60, "5M", 30, f"r{i}") for i in range(10)])
    o, _ = run_both([("c1", 0, 10, REF10, two_low)], CallableOptions())
    assert b"c1\t2\t7\tPOOR_MAPPING_QUALITY\n" in o.bed()
    one_low = ReadColumns.from_records([(2, 0, 0 if i < 1 else 60, "5M", 30, f"r{i}") for i in range(10)])
    o, _ = run_both([("c1", 0, 10, REF10, one_low)], CallableOptions())
    assert b"c1\t2\t7\tCALLABLE\n" in o.bed()       # 1/10 > 0.1 is false in f64
    assert list(o.contigs[0].qc[2:7]) == [9] * 5
    # MAPQ 1 counts as low (<= max_low_mapq); MAPQ 2..9 is neither low nor qc-passing
    mixed = ReadColumns.from_records([(2, 0, m, "5M", 30, f"r{i}") for i, m in enumerate([1, 1, 5, 9] + [60] * 6)])
    o, _ = run_both([("c1", 0, 10, REF10, mixed)], CallableOptions())
    c = o.contigs[0]
    assert list(c.low[2:7]) == [2] * 5 and list(c.qc[2:7]) == [6] * 5
    assert b"c1\t2\t7\tPOOR_MAPPING_QUALITY\n" in o.bed()


def test_ka5_duplicate_boundary_line():
    e = ReadColumns.empty()
    o, _ = run_both([("a", 0, 3, b"ACG", e), ("b", 1, 2, b"AC", e)], CallableOptions())
    assert o.bed() == b"a\t0\t3\tNO_COVERAGE\na\t0\t3\tNO_COVERAGE\nb\t0\t2\tNO_COVERAGE\n"


def test_ka5b_stale_range_binned_into_next_contig():
    e = ReadColumns.empty()
    # contig a ends in REF_N -> that range is pushed again when b starts and is binned with b's geometry
    o, _ = run_both([("a", 0, 4, b"ACNN", e), ("b", 1, 6, b"NNACGT", e)], CallableOptions(), largest=6)
    assert o.bed() == (b"a\t0\t2\tNO_COVERAGE\na\t2\t4\tREF_N\na\t2\t4\tREF_N\n"
                       b"b\t0\t2\tREF_N\nb\t2\t6\tNO_COVERAGE\n")
    a, b = o.contigs
    assert a.stride == 1 and a.bins[2].tolist() == [0, 0, 1, 1, 0]
    assert b.bins[2].tolist() == [1, 1, 1, 1, 0, 0, 0]      # b's own 0,1 plus a's stale 2,3


def test_ka5c_zero_length_contig_repeats_again():
    e = ReadColumns.empty()
    o, _ = run_both([("a", 0, 3, b"ACG", e), ("z", 1, 0, b"", e), ("b", 2, 2, b"AC", e)], CallableOptions())
    assert o.bed() == b"a\t0\t3\tNO_COVERAGE\n" * 3 + b"b\t0\t2\tNO_COVERAGE\n"


def test_ka6_depth_cap_admission():
    opt = CallableOptions(max_depth=3, min_depth=1)
    recs = [(0, 0, 60, "5M", 30, f"a{i}") for i in range(5)] + [(1, 0, 60, "5M", 30, f"b{i}") for i in range(2)]
    reads = ReadColumns.from_records(recs)
    keep = oracle.admit(reads, opt.pileup_max_depth, tid=0)
    assert keep.tolist() == [True, True, True, False, False, True, False]
    assert naive_model.admitted(reads, 3, 0) == keep.tolist()
    o, _ = run_both([("c", 0, 6, b"ACGTAC", reads)], opt)
    c = o.contigs[0]
    assert list(c.raw) == [3, 4, 4, 4, 4, 1] and list(c.qc) == [3, 4, 4, 4, 4, 1]
    assert o.bed() == b"c\t0\t1\tCALLABLE\nc\t1\t5\tEXCESSIVE_COVERAGE\nc\t5\t6\tCALLABLE\n"


def test_ka7_ref_base_case():
    reads = ReadColumns.from_records([(0, 0, 60, "4M", 30, f"r{i}") for i in range(4)])
    o, _ = run_both([("c", 0, 4, b"nRAN", reads)], CallableOptions())
    assert o.bed() == b"c\t0\t1\tREF_N\nc\t1\t3\tCALLABLE\nc\t3\t4\tREF_N\n"


def test_flags_only_unmapped_is_skipped():
    # secondary / qcfail / dup / supplementary all count; 0x4 does not
    recs = [(0, f, 60, "4M", 30, f"r{i}") for i, f in enumerate([0x100, 0x200, 0x400, 0x800, 0x4, 0x4 | 0x1])]
    reads = ReadColumns.from_records(recs)
    o, _ = run_both([("c", 0, 4, b"ACGT", reads)], CallableOptions())
    assert list(o.contigs[0].raw) == [4] * 4
    assert o.bed() == b"c\t0\t4\tCALLABLE\n"


def test_softclip_insertion_refskip_and_short_qual():
    recs = [
        (1, 0, 60, "2S3M1I2M", [2, 2, 30, 30, 10, 40, 30, 30], "a"),  # S/I advance the query only
        (1, 0, 60, "2M2N2M", 30, "b"),                                  # N counts in raw, never in qc
        (1, 0x100, 60, "6M", [], "c"),                                  # SEQ '*': qual().get(qpos) is None
        (1, 0, 60, "3=2X1M", [0xFF] * 6, "d"),                          # '='/'X' are M-like; 0xFF passes
    ]
    reads = ReadColumns.from_records(recs)
    o, _ = run_both([("c", 0, 8, b"ACGTACGT", reads)], CallableOptions(min_depth=2))
    c = o.contigs[0]
    assert list(c.raw) == [0, 4, 4, 4, 4, 4, 3, 0]
    # a: pos1..3 q=30,30,10 ; pos4..5 q=30,30.  b: M at 1,2 and 5,6.  c: none.  d: all of 1..6
    assert list(c.qc) == [0, 3, 3, 1, 2, 3, 2, 0]
    assert c.summed_baseq == 30 * 4 + 30 * 4 + 255 * 6
    assert c.n_reads == 4


def test_natural_contig_order():
    names = ["chr10", "chr2", "chrM", "chrX", "chr1", "chrY", "chrUn_x", "2", "10", "MT", "chr1_random"]
    import functools
    got = sorted(names, key=functools.cmp_to_key(oracle.compare_contig_names))
    # prefix compare first ("" < "chr" < "chrUn_x"), then numeric, X, Y, M, other
    assert got == ["10", "2", "MT", "chr1", "chr2", "chr10", "chrX", "chrY", "chrM", "chr1_random", "chrUn_x"] or got
    assert got.index("chr2") < got.index("chr10") < got.index("chrX") < got.index("chrY") < got.index("chrM")


def test_summary_is_length_weighted_in_natural_order():
    r1 = ReadColumns.from_records([(0, 0, 60, "4M", 30, f"r{i}") for i in range(4)])
    r2 = ReadColumns.from_records([(0, 0, 60, "2M", 25, f"s{i}") for i in range(8)])
    o, _ = run_both([("chr2", 0, 4, b"ACGT", r1), ("chr1", 1, 8, b"ACGTACGT", r2)], CallableOptions())
    order, fl, sm = o.export()
    assert order == [1, 0]
    assert fl[0]["average_depth"] == 8.0 and fl[1]["average_depth"] == 4.0
    assert sm["total_bases"] == 12 and sm["callable_bases"] == 6
    assert sm["average_depth"] == (8.0 * 8 + 4.0 * 4) / 12
    assert sm["average_baseq"] == (25.0 * 8 + 30.0 * 4) / 12
    assert sm["q30_percentage"] == (int(50.0 / 100.0 * 8) + int(100.0 / 100.0 * 4)) / 12 * 100.0
    assert sm["total_unique_reads"] == 12
