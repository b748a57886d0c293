"""GPU parity tests proper: the CUDA path, called through the C-ABI, against the CPU oracle on the same
seeded inputs.  Bar: BED bytes identical, every integer bit-exact, floats equal."""
import numpy as np
import pytest

from decodingustools_b200 import synth
from decodingustools_b200.callable_loci import (CallableLociContext, admit_reads, compact_reads, stitch_intervals)
from decodingustools_b200.options import CallableOptions
from decodingustools_b200.soa import ReadColumns
from tests.helpers import assert_parity, run_oracle
from tests.test_host_half import states_to_intervals
from tests.test_oracle_vs_naive import random_reads

pytestmark = pytest.mark.gpu

from decodingustools_b200 import _lib
WREAL = int(_lib.lib().clb_window_positions())
REF10 = b"NNACGTACGT"


@pytest.fixture(scope="module")
def ctx_default():
    c = CallableLociContext(CallableOptions())
    yield c
    c.close()


def test_known_answers_through_the_device(ctx_default):
    opt = CallableOptions()
    four = ReadColumns.from_records([(2, 0, 60, "5M", 30, f"r{i}") for i in range(4)])
    dele = ReadColumns.from_records([(2, 0, 60, "2M1D2M", 30, f"r{i}") for i in range(4)])
    two_low = ReadColumns.from_records([(2, 0, 0 if i < 2 else 60, "5M", 30, f"r{i}") for i in range(10)])
    one_low = ReadColumns.from_records([(2, 0, 0 if i < 1 else 60, "5M", 30, f"r{i}") for i in range(10)])
    for reads, expect in [(ReadColumns.empty(), b"c1\t0\t2\tREF_N\nc1\t2\t10\tNO_COVERAGE\n"),
                          (four, b"c1\t0\t2\tREF_N\nc1\t2\t7\tCALLABLE\nc1\t7\t10\tNO_COVERAGE\n"),
                          (dele, b"c1\t0\t2\tREF_N\nc1\t2\t4\tCALLABLE\nc1\t4\t5\tLOW_COVERAGE\nc1\t5\t7\tCALLABLE\nc1\t7\t10\tNO_COVERAGE\n"),
                          (two_low, b"c1\t0\t2\tREF_N\nc1\t2\t7\tPOOR_MAPPING_QUALITY\nc1\t7\t10\tNO_COVERAGE\n"),
                          (one_low, b"c1\t0\t2\tREF_N\nc1\t2\t7\tCALLABLE\nc1\t7\t10\tNO_COVERAGE\n")]:
        o, _ = assert_parity([("c1", 0, 10, REF10, reads)], opt, ctx_default)
        assert o.bed() == expect


def test_ka5_ka6_quirks_and_depth_cap():
    e = ReadColumns.empty()
    o, _ = assert_parity([("a", 0, 3, b"ACG", e), ("z", 1, 0, b"", e), ("b", 2, 2, b"AC", e)], CallableOptions())
    assert o.bed() == b"a\t0\t3\tNO_COVERAGE\n" * 3 + b"b\t0\t2\tNO_COVERAGE\n"
    opt = CallableOptions(max_depth=3, min_depth=1)
    recs = [(0, 0, 60, "5M", 30, f"a{i}") for i in range(5)] + [(1, 0, 60, "5M", 30, f"b{i}") for i in range(2)]
    o, _ = assert_parity([("c", 0, 6, b"ACGTAC", ReadColumns.from_records(recs))], opt)
    assert o.bed() == b"c\t0\t1\tCALLABLE\nc\t1\t5\tEXCESSIVE_COVERAGE\nc\t5\t6\tCALLABLE\n"


@pytest.mark.parametrize("seed", range(25))
def test_random_micro_contigs(seed):
    rng = np.random.default_rng(1000 + seed)
    length = int(rng.integers(1, 60))
    ref = bytes(rng.choice(list(b"ACGTNnR"), size=length, p=[.2, .2, .2, .2, .1, .05, .05]).tolist())
    reads = random_reads(rng, length, int(rng.integers(0, 80)))
    opt = CallableOptions(min_depth=int(rng.integers(0, 5)), max_depth=int(rng.choice([0, 2, 3, 5, 500])),
                          min_mapping_quality=int(rng.choice([0, 10, 30])), min_base_quality=int(rng.choice([0, 20, 200])),
                          min_depth_for_low_mapq=int(rng.integers(0, 6)), max_low_mapq=int(rng.choice([0, 1, 9])),
                          max_low_mapq_fraction=float(rng.choice([-0.5, 0.0, 0.1, 0.25, 0.5, 1.0])))
    assert_parity([("chrT", int(rng.integers(0, 3)), length, ref, reads)], opt)


@pytest.mark.parametrize("seed", range(6))
def test_random_multi_contig_quirks(seed, ctx_default):
    rng = np.random.default_rng(5000 + seed)
    contigs = []
    for tid, name in enumerate(["chr1", "chr2", "chrM", "chrX", "chrUn"]):
        length = int(rng.choice([0, 1, 5, 17, 40, 5000]))
        ref = bytes(rng.choice(list(b"ACGTN"), size=length, p=[.22, .22, .22, .22, .12]).tolist())
        reads = random_reads(rng, max(length, 1), int(rng.integers(0, 200)), max_len=30) if length > 2 else ReadColumns.empty()
        contigs.append((name, tid, length, ref, reads))
    if max(c[2] for c in contigs if c[0] != "chrM") == 0:
        contigs[0] = ("chr1", 0, 5, b"ACGTN", ReadColumns.empty())
    assert_parity(contigs, CallableOptions(), ctx_default)


@pytest.mark.parametrize("length", [1, 2, WREAL - 1, WREAL, WREAL + 1, 2 * WREAL, 2 * WREAL + 1, 3 * WREAL - 1])
def test_window_edges(length, ctx_default):
    """Contig lengths and reads placed exactly on window seams."""
    rng = np.random.default_rng(length)
    ref = bytes(rng.choice(list(b"ACGTN"), size=length, p=[.24, .24, .24, .24, .04]).tolist())
    recs = []
    for seam in (0, WREAL, 2 * WREAL):
        for d in (-151, -150, -149, -2, -1, 0, 1):
            p = seam + d
            if 0 <= p and p + 150 <= length:
                for k in range(5):
                    recs.append((p, 0, int(rng.choice([0, 60])), "150M", rng.choice([2, 37], size=150).tolist(), f"s{seam}_{d}_{k}"))
                recs.append((p, 0, 60, "20S50M10D50M30S", rng.choice([2, 37], size=150).tolist(), f"c{seam}_{d}"))
    recs.sort(key=lambda r: r[0])
    assert_parity([("chrE", 0, length, ref, ReadColumns.from_records(recs))], CallableOptions(), ctx_default)


def test_synthetic_short_reads_per_base_and_bed(ctx_default):
    c = synth.synth_short("chr22", 1_500_000, seed=11)
    opt = CallableOptions()
    contigs = [(c.name, 0, c.length, c.ref, c.reads)]
    o, results = assert_parity(contigs, opt, ctx_default)
    # per-base counters of the same run
    oc = run_oracle(contigs, opt, debug=True).contigs[0]
    raw, qc, low, st = ctx_default.debug_per_base(c.length)
    assert np.array_equal(raw, oc.raw) and np.array_equal(qc, oc.qc) and np.array_equal(low, oc.low)
    assert np.array_equal(st, oc.state)
    assert results[0].summed_coverage == int(c.reads.select(admit_reads(c.reads, 500)).ref_len().sum())
    # the per-base dump comes from the kernels' debug instantiations (<.., DBG = true>): what they left on the device must be
    # exactly what the production instantiations produced for the same resident contig
    dbg = ctx_default.refresh_counters()
    prod = results[0]
    assert np.array_equal(dbg.intervals, prod.intervals) and np.array_equal(dbg.state_counts, prod.state_counts) and np.array_equal(dbg.bins, prod.bins)
    for k in ("n_covered_bases", "summed_coverage", "summed_baseq", "summed_mapq", "quality_bases"):
        assert getattr(dbg, k) == getattr(prod, k), k


def test_streamed_batches_equal_single_push(ctx_default):
    c = synth.synth_short("chr22", 600_000, seed=12)
    assert_parity([(c.name, 0, c.length, c.ref, c.reads)], CallableOptions(), ctx_default, batch_reads=10_000)
    assert_parity([(c.name, 0, c.length, c.ref, c.reads)], CallableOptions(), ctx_default, batch_reads=777)


def test_long_reads_indel_heavy(ctx_default):
    c = synth.synth_long("chr1", 600_000, seed=13)
    assert c.reads.n_cigar > 50 * c.reads.n
    assert_parity([(c.name, 0, c.length, c.ref, c.reads)], CallableOptions(), ctx_default)
    assert_parity([(c.name, 0, c.length, c.ref, c.reads)], CallableOptions(), ctx_default, batch_reads=7)


def test_deep_coverage_admission_cap():
    c = synth.synth_short("chrM", 16_569, seed=14, depth=2000.0)
    y = synth.synth_short("chrY", 120_000, seed=15, depth=2000.0)
    o, _ = assert_parity([(y.name, 0, y.length, y.ref, y.reads), (c.name, 1, c.length, c.ref, c.reads)], CallableOptions())
    assert o.contigs[0].n_admitted < y.reads.n // 2            # the cap is active
    # max_depth 4000: no cap, true 2000x depth in the shared-memory counters
    assert_parity([(c.name, 0, c.length, c.ref, c.reads)], CallableOptions(max_depth=4000))


def test_region_shards_stitch_to_the_whole(ctx_default):
    c = synth.synth_short("chr22", 500_000, seed=16)
    opt = CallableOptions()
    keep = admit_reads(c.reads, opt.pileup_max_depth, 0)
    reads = c.reads.select(keep)
    span = reads.max_ref_span()
    ctx = ctx_default
    ctx.begin_contig(0, c.name, c.length, c.ref, c.length, max_ref_span=span)
    ctx.push_reads(reads)
    whole = ctx.finish_contig()
    for cuts in ([0, 250_000, c.length], [0, WREAL, WREAL + 1, 123_457, 400_000, c.length]):
        parts = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            end = reads.end()
            sel = (reads.pos < b) & (end > a - 1)               # halo: the base left of the shard too
            ctx.begin_contig(0, c.name, c.length, c.ref, c.length, region=(a, b), max_ref_span=span)
            ctx.push_reads(reads.select(sel))
            parts.append(ctx.finish_contig())
        iv = stitch_intervals([p.intervals for p in parts])
        assert np.array_equal(iv, whole.intervals)
        assert np.array_equal(sum(p.state_counts for p in parts), whole.state_counts)
        assert np.array_equal(sum(p.bins.astype(np.uint64) for p in parts), whole.bins.astype(np.uint64))
        for k in ("n_covered_bases", "summed_coverage", "summed_baseq", "summed_mapq", "quality_bases"):
            assert sum(getattr(p, k) for p in parts) == getattr(whole, k), k


def test_unsorted_input_is_rejected(ctx_default):
    from decodingustools_b200._lib import ClbError
    reads = ReadColumns.from_records([(50, 0, 60, "10M", 30), (20, 0, 60, "10M", 30)])
    ctx_default.begin_contig(0, "c", 100, b"A" * 100, 100)
    ctx_default.push_reads(reads)
    with pytest.raises(ClbError):
        ctx_default.finish_contig()


def test_rerun_resident_is_idempotent(ctx_default):
    c = synth.synth_short("chr22", 400_000, seed=17)
    opt = CallableOptions()
    reads = c.reads.select(admit_reads(c.reads, 500, 0))
    ctx_default.begin_contig(0, c.name, c.length, c.ref, c.length, max_ref_span=reads.max_ref_span())
    ctx_default.push_reads(reads)
    first = ctx_default.finish_contig()
    for _ in range(3):
        ms, again = ctx_default.rerun_resident(fetch=True, copy_intervals=True)
        assert ms > 0
        assert np.array_equal(again.intervals, first.intervals) and np.array_equal(again.state_counts, first.state_counts)
        assert again.summed_baseq == first.summed_baseq and np.array_equal(again.bins, first.bins)


def test_whole_genome_layout_tiny_scale(ctx_default):
    """BASELINE config 3 shape: 24 contigs + chrM in tid order, quirks Q1/Q2 across every contig boundary."""
    contigs = synth.config(2, scale=0.0015)
    assert len(contigs) == 25 and contigs[-1].name == "chrM"
    assert_parity([(c.name, tid, c.length, c.ref, c.reads) for tid, c in enumerate(contigs)], CallableOptions(), ctx_default)


def test_device_computed_span_and_nmask_reference(ctx_default):
    """max_ref_span = 0 lets the library compute the look-back on the device; the reference may come as a bit-packed N mask."""
    from decodingustools_b200.callable_loci import compact_reads
    from decodingustools_b200.soa import n_mask_from_ascii
    c = synth.synth_short("chr22", 300_000, seed=31)
    reads = compact_reads(c.reads, admit_reads(c.reads, 500, 0))
    ctx_default.begin_contig(0, c.name, c.length, c.ref, c.length, max_ref_span=reads.max_ref_span())
    ctx_default.push_reads(reads)
    want = ctx_default.finish_contig()
    ctx_default.begin_contig(0, c.name, c.length, n_mask_from_ascii(c.ref), c.length, max_ref_span=0, ref_is_nmask=True)
    for lo in range(0, reads.n, 20_000):
        ctx_default.push_reads(reads.slice(lo, lo + 20_000))
    got = ctx_default.finish_contig()
    assert np.array_equal(got.intervals, want.intervals) and np.array_equal(got.state_counts, want.state_counts)
    assert np.array_equal(got.bins, want.bins) and got.summed_baseq == want.summed_baseq and got.summed_mapq == want.summed_mapq


def test_reference_shorter_than_contig_reads_as_n(ctx_default):
    """fetch_seq past the FASTA end yields nothing -> 'N' (mod.rs:79-80)."""
    reads = ReadColumns.from_records([(p, 0, 60, "50M", 30, f"r{p}_{k}") for p in range(0, 150, 10) for k in range(5)])
    assert_parity([("c", 0, 300, b"ACGT" * 30, reads)], CallableOptions(), ctx_default)


def test_intervals_tile_the_contig_and_alternate(ctx_default):
    """Size-independent properties used at full size by bench.py."""
    c = synth.synth_short("chr22", 2_000_000, seed=33)
    from decodingustools_b200.callable_loci import compact_reads
    reads = compact_reads(c.reads, admit_reads(c.reads, 500, 0))
    ctx_default.begin_contig(0, c.name, c.length, c.ref, c.length, max_ref_span=reads.max_ref_span())
    ctx_default.push_reads(reads)
    r = ctx_default.finish_contig()
    iv = r.intervals
    assert iv["start"][0] == 0 and iv["end"][-1] == c.length
    assert np.array_equal(iv["start"][1:], iv["end"][:-1]) and np.all(iv["state"][1:] != iv["state"][:-1])
    assert int(r.state_counts.sum()) == c.length
    lens = (iv["end"] - iv["start"]).astype(np.int64)
    assert np.array_equal(np.bincount(iv["state"], weights=lens, minlength=6).astype(np.int64), r.state_counts.astype(np.int64))
    assert r.summed_coverage == int(reads.ref_len().sum())
    assert r.n_covered_bases >= int(r.state_counts[1] + r.state_counts[3] + r.state_counts[4] + r.state_counts[5])


def _pile(rng, n, length, same_start, with_indels):
    """n reads over a short contig: random or identical starts, 1-3 ops, random MAPQ / qualities / flags."""
    pos = np.zeros(n, np.int64) if same_start else np.sort(rng.integers(0, max(1, length - 70), n))
    m1 = rng.integers(20, 40, n)
    three = (rng.random(n) < 0.3) if with_indels else np.zeros(n, bool)
    gap = rng.integers(1, 5, n); m2 = rng.integers(5, 25, n)
    is_del = rng.random(n) < 0.5
    nops = np.where(three, 3, 1)
    cigar_off = np.concatenate([[0], np.cumsum(nops)])
    cigar = np.zeros(int(cigar_off[-1]), np.uint32)
    cigar[cigar_off[:-1]] = (m1 << 4) | 0
    t = np.nonzero(three)[0]
    cigar[cigar_off[t] + 1] = (gap[t] << 4) | np.where(is_del[t], 2, 1)
    cigar[cigar_off[t] + 2] = (m2[t] << 4) | 0
    qlen = m1 + np.where(three, m2 + np.where(is_del, 0, gap), 0)
    qual_off = np.concatenate([[0], np.cumsum(qlen)])
    qual = rng.integers(2, 42, int(qual_off[-1])).astype(np.uint8)
    mapq = rng.choice(np.array([0, 1, 5, 30, 60], np.uint8), n, p=[.2, .1, .1, .2, .4])
    flag = rng.choice(np.array([0, 16, 1024, 4], np.uint16), n, p=[.5, .4, .08, .02])
    rc = ReadColumns(pos, flag, mapq, cigar_off, cigar, qual_off, qual, np.arange(n, dtype=np.uint32))
    if not with_indels:
        return rc
    # every 50th read gets a 9-op CIGAR (the warp-cooperative long-CIGAR path)
    recs = []
    for i in range(n):
        if i % 50 == 0:
            ops = [(int(rng.integers(3, 9)), 0)]
            for _ in range(4):
                ops += [(int(rng.integers(1, 4)), int(rng.choice([1, 2, 3, 4]))), (int(rng.integers(3, 9)), 0)]
            ql = sum(l for l, o in ops if o in (0, 1, 4))
            recs.append((int(pos[i]), int(flag[i]), int(mapq[i]), "".join(f"{l}{'MIDNS'[o]}" for l, o in ops), rng.integers(2, 42, ql).astype(np.uint8).tobytes(), f"l{i}"))
    longs = ReadColumns.from_records(recs)
    keep = np.ones(n, bool); keep[::50] = False
    short = rc.select(keep)
    order = np.argsort(np.concatenate([short.pos, longs.pos]), kind="stable")
    both = _concat(short, longs)
    return _reorder(both, order)


def _concat(a, b):
    return ReadColumns(np.concatenate([a.pos, b.pos]), np.concatenate([a.flag, b.flag]), np.concatenate([a.mapq, b.mapq]),
                       np.concatenate([a.cigar_off, b.cigar_off[1:] + a.cigar_off[-1]]), np.concatenate([a.cigar, b.cigar]),
                       np.concatenate([a.qual_off, b.qual_off[1:] + a.qual_off[-1]]), np.concatenate([a.qual, b.qual]),
                       np.arange(a.n + b.n, dtype=np.uint32))


def _reorder(rc, order):
    """Records of rc in the given order (used to merge two coordinate-sorted sets)."""
    co, qo = rc.cigar_off.astype(np.int64), rc.qual_off.astype(np.int64)
    cl, ql = np.diff(co)[order], np.diff(qo)[order]
    cigar = np.concatenate([rc.cigar[co[i]:co[i + 1]] for i in order]) if len(order) else rc.cigar
    qual = np.concatenate([rc.qual[qo[i]:qo[i + 1]] for i in order]) if len(order) else rc.qual
    return ReadColumns(rc.pos[order], rc.flag[order], rc.mapq[order], np.concatenate([[0], np.cumsum(cl)]), cigar,
                       np.concatenate([[0], np.cumsum(ql)]), qual, np.arange(len(order), dtype=np.uint32))


@pytest.mark.parametrize("same_start", [False, True])
def test_more_than_65535_reads_in_one_window(same_start):
    """Windows whose candidate count does not fit the packed 16-bit depth fields take the second (32-bit) pass; with
    identical starts the raw depth itself passes 65535 and the low-MAPQ threshold is computed past the lookup table."""
    rng = np.random.default_rng(4242 + int(same_start))
    length = 2 * WREAL + 300                                   # deep windows next to an ordinary, shallow one
    reads = _pile(rng, 70_000, WREAL - 100, same_start, with_indels=not same_start)
    tail = _pile(rng, 300, 250, False, True)
    tail.pos += 2 * WREAL
    both = _concat(reads, tail)
    ref = bytes(rng.choice(list(b"ACGTN"), size=length, p=[.245, .245, .245, .245, .02]).tolist())
    for opt in (CallableOptions(max_depth=100_000, max_low_mapq_fraction=0.3), CallableOptions(max_depth=0, min_depth=1, max_low_mapq=5)):
        o, _ = assert_parity([("chrD", 0, length, ref, both)], opt)
        if opt.max_depth:
            assert o.contigs[0].n_admitted > 65_535 and (not same_start or max(o.contigs[0].counts) > 0)


def test_more_runs_than_the_first_record_buffer_holds(ctx_default):
    """A reference that alternates N / non-N at every base yields one BED run per base: far more than the initial
    boundary-record capacity (a quarter of the region), so the contig is re-run with a grown buffer."""
    length = 400_001
    ref = (b"AN" * (length // 2 + 1))[:length]
    reads = ReadColumns.from_records([(p, 0, 60, "100M", 35, f"r{p}_{k}") for p in range(1000, 200_000, 37) for k in range(2)])
    o, g = assert_parity([("alt", 0, length, ref, reads)], CallableOptions(), ctx_default)
    assert o.bed().count(b"\n") == length                                   # every run is one base long
    # the grown buffer is kept: a second contig through the same context still agrees
    assert_parity([("alt2", 0, 70_000, ref[:70_000], ReadColumns.empty())], CallableOptions(), ctx_default)


def test_reads_past_the_contig_end_are_clipped_not_refused(ctx_default):
    """include/callable_loci_b200.h: a read that runs past the contig end is clipped there (the reference has no guard and
    would walk past it): the result equals the one of the same reads cut at the end by hand."""
    ref = b"ACGT" * 25
    long_recs = [(90, 0, 60, "20M", list(range(20, 40)), f"a{i}") for i in range(5)] + [(95, 0, 60, "3M4D10M", 30, "d")]
    cut_recs = [(90, 0, 60, "10M", list(range(20, 30)), f"a{i}") for i in range(5)] + [(95, 0, 60, "3M2D", 30, "d")]
    res = []
    for recs in (long_recs, cut_recs):
        rc = ReadColumns.from_records(recs)
        ctx_default.begin_contig(0, "c", 100, ref, 100, max_ref_span=rc.max_ref_span())
        ctx_default.push_reads(rc)
        res.append(ctx_default.finish_contig())
    a, b = res
    assert np.array_equal(a.intervals, b.intervals) and np.array_equal(a.state_counts, b.state_counts)
    for k in ("n_covered_bases", "summed_coverage", "summed_baseq", "summed_mapq", "quality_bases"):
        assert getattr(a, k) == getattr(b, k), k
    assert a.summed_coverage == 5 * 10 + 5


def test_malformed_batches_are_refused_with_input_errors(ctx_default):
    from decodingustools_b200._lib import ClbError
    pos = np.array([5, 7, 9], np.int32); flag = np.zeros(3, np.uint16); mapq = np.full(3, 60, np.uint8)
    cig = np.array([160, 160, 160], np.uint32); qoff = np.array([0, 10, 20, 30], np.uint64); qual = np.full(30, 30, np.uint8)
    keep = [pos, flag, mapq, cig, qoff, qual]
    p = lambda a: a.ctypes.data

    def push(coff):
        keep.append(coff)
        ctx_default.begin_contig(0, "c", 100, b"A" * 100, 100)
        ctx_default.push_raw(3, 3, 30, p(pos), p(flag), p(mapq), p(coff), p(cig), p(qoff), p(qual))
    with pytest.raises(ClbError):
        push(np.array([0, 1, 2, 2], np.uint32))              # offsets do not end at n_cigar: refused before anything is copied
    with pytest.raises(ClbError):
        push(np.array([1, 1, 2, 3], np.uint32))              # offsets do not start at 0
    push(np.array([0, 2, 1, 3], np.uint32))                  # ends are right, the middle is not monotone: caught on the device, kernels do not walk it
    with pytest.raises(ClbError):
        ctx_default.finish_contig()
    ctx_default.begin_contig(0, "c", 100, b"A" * 100, 100)   # the context is usable again afterwards
    ok_off = np.array([0, 1, 2, 3], np.uint32)
    ctx_default.push_raw(3, 3, 30, p(pos), p(flag), p(mapq), p(ok_off), p(cig), p(qoff), p(qual))
    assert ctx_default.finish_contig().summed_coverage == 30


@pytest.mark.parametrize("min_depth,max_depth,min_dflm,frac", [
    (0, 0, 0, 0.0), (4, 127, 10, 0.1), (127, 128, 128, 0.0), (128, 253, 120, 0.5), (160, 254, 300, 0.1), (200, 255, 0, -0.5),
    (255, 1000, 5, 1.0), (256, 100, 129, 0.25), (1000, 1, 200, 0.0), (100, 140, 100, 0.02)])
def test_byte_wide_thresholds_on_a_deep_short_read_pile(min_depth, max_depth, min_dflm, frac):
    """Depths of 100-200 with thresholds on both sides of 128 / 255: every depth of a fast window is held in byte-wide counters."""
    c = synth.synth_short("chr21", 40_000, seed=77, depth=150.0)
    opt = CallableOptions(min_depth=min_depth, max_depth=max_depth, min_depth_for_low_mapq=min_dflm, max_low_mapq_fraction=frac,
                          max_low_mapq=3)
    o, results = assert_parity([(c.name, 0, c.length, c.ref, c.reads)], opt)
    assert results[0].general_windows < 5          # the pile stays below 255: the fast kernel takes (nearly) all windows


def test_long_reads_arriving_after_short_read_batches(ctx_default):
    """A contig whose first batches are short reads and whose later batches are indel-heavy long reads: long-read mode
    (checkpoints, general kernel) switches on at the first long batch and the result is the oracle's either way."""
    a = synth.synth_short("chr5", 120_000, seed=21)
    b = synth.synth_long("chr5", 120_000, seed=22, depth=8.0)
    first = a.reads.select(a.reads.pos < 60_000)
    later = b.reads.select(b.reads.pos >= 60_000)
    reads = ReadColumns(np.concatenate([first.pos, later.pos]), np.concatenate([first.flag, later.flag]),
                        np.concatenate([first.mapq, later.mapq]),
                        np.concatenate([first.cigar_off, later.cigar_off[1:] + first.cigar_off[-1]]).astype(np.uint32),
                        np.concatenate([first.cigar, later.cigar]),
                        np.concatenate([first.qual_off, later.qual_off[1:] + first.qual_off[-1]]).astype(np.uint64),
                        np.concatenate([first.qual, later.qual]), np.arange(first.n + later.n, dtype=np.uint32))
    assert np.all(np.diff(reads.pos.astype(np.int64)) >= 0)
    for batch_reads in (0, first.n, 5000):
        assert_parity([("chr5", 0, a.length, a.ref, reads)], CallableOptions(), ctx_default, batch_reads=batch_reads)


def test_queued_reruns_leave_the_same_result(ctx_default):
    """clb_rerun_resident(ctx, NULL, NULL) only enqueues: three steps queued back to back, waited for once."""
    c = synth.synth_short("chr22", 400_000, seed=31)
    ctx_default.begin_contig(0, c.name, c.length, c.ref, c.length, max_ref_span=c.reads.max_ref_span())
    ctx_default.push_reads(compact_reads(c.reads, admit_reads(c.reads, 500)))
    first = ctx_default.finish_contig()
    for _ in range(3):
        assert ctx_default.rerun_resident(fetch=False, sync=False) == (None, None)
    again = ctx_default.refresh_counters()
    assert np.array_equal(again.intervals, first.intervals) and np.array_equal(again.state_counts, first.state_counts)
    assert np.array_equal(again.bins, first.bins) and again.summed_baseq == first.summed_baseq and again.summed_mapq == first.summed_mapq
