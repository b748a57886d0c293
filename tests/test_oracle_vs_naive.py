"""Randomised micro-contigs: C oracle (htslib iterator mechanics) vs naive Python model (closed-form rules)."""
import numpy as np
import pytest

from decodingustools_b200.options import CallableOptions
from decodingustools_b200.soa import ReadColumns
from oracle import naive_model, oracle
from tests.test_oracle_known_answers import run_both


def random_reads(rng, length, n, max_len=12, zero_span=True):
    recs = []
    pos = np.sort(rng.integers(0, max(1, length - 1), size=n))
    for i, p in enumerate(pos):
        kind = rng.integers(0, 10)
        room = max(1, length - int(p))
        L = int(rng.integers(1, min(max_len, room) + 1))
        if kind == 0 and L >= 3:
            a = int(rng.integers(1, L - 1)); d = int(rng.integers(1, L - a)); b = L - a - d
            cig = f"{a}M{d}D{b}M" if b > 0 else f"{a}M{d}D"      # may end in a deletion on purpose
        elif kind == 1 and L >= 2:
            a = int(rng.integers(1, L)); cig = f"{a}M{int(rng.integers(1, 4))}I{L - a}M"
        elif kind == 2:
            cig = f"{int(rng.integers(1, 4))}S{L}M{int(rng.integers(0, 3))}S".replace("0S", "")
        elif kind == 3 and L >= 3:
            a = int(rng.integers(1, L - 1)); d = int(rng.integers(1, L - a)); b = L - a - d
            cig = f"{a}M{d}N{b}M" if b > 0 else f"{L}M"
        elif kind == 4 and zero_span:
            cig = rng.choice(["*", "3S", "2I", "2H3S"])
        elif kind == 5:
            cig = f"2H{L}={1}X" if L + 1 <= room else f"{L}="
        else:
            cig = f"{L}M"
        ops = [(int(x[:-1]), x[-1]) for x in __import__("re").findall(r"\d+[MIDNSHP=X]", cig)]
        qlen = sum(n_ for n_, o in ops if o in "MIS=X")
        if rng.integers(0, 12) == 0:
            q = []                                     # SEQ '*'
        else:
            q = rng.choice([2, 12, 19, 20, 23, 37, 255], size=qlen).tolist()
        mapq = int(rng.choice([0, 1, 2, 9, 10, 30, 60]))
        flag = int(rng.choice([0, 0, 0, 0x10, 0x100, 0x400, 0x800, 0x4, 0x200]))
        recs.append((int(p), flag, mapq, cig, q, f"q{int(rng.integers(0, max(1, n // 2)))}"))
    return ReadColumns.from_records(recs)


@pytest.mark.parametrize("seed", range(40))
def test_random_single_contig(seed):
    rng = np.random.default_rng(1000 + seed)
    length = int(rng.integers(1, 60))
    ref = bytes(rng.choice(list(b"ACGTNnR"), size=length, p=[.2, .2, .2, .2, .1, .05, .05]).tolist())
    reads = random_reads(rng, length, int(rng.integers(0, 80)))
    opt = CallableOptions(min_depth=int(rng.integers(0, 5)), max_depth=int(rng.choice([0, 2, 3, 5, 500])),
                          min_mapping_quality=int(rng.choice([0, 10, 30])), min_base_quality=int(rng.choice([0, 20, 200])),
                          min_depth_for_low_mapq=int(rng.integers(0, 6)), max_low_mapq=int(rng.choice([0, 1, 9])),
                          max_low_mapq_fraction=float(rng.choice([0.0, 0.1, 0.25, 0.5, 1.0])))
    tid = int(rng.integers(0, 3))
    run_both([("chrT", tid, length, ref, reads)], opt)
    keep = oracle.admit(reads, opt.pileup_max_depth, tid=tid)
    assert keep.tolist() == naive_model.admitted(reads, opt.pileup_max_depth, tid)


@pytest.mark.parametrize("seed", range(10))
def test_random_multi_contig_with_quirks(seed):
    rng = np.random.default_rng(5000 + seed)
    names = ["chr1", "chr2", "chrM", "chrX", "chrUn"]
    contigs = []
    for tid, name in enumerate(names):
        length = int(rng.choice([0, 1, 5, 17, 40]))
        ref = bytes(rng.choice(list(b"ACGTN"), size=length, p=[.22, .22, .22, .22, .12]).tolist())
        reads = random_reads(rng, max(length, 1), int(rng.integers(0, 40))) if length > 2 else ReadColumns.empty()
        contigs.append((name, tid, length, ref, reads))
    if max(c[2] for c in contigs if c[0] != "chrM") == 0:
        contigs[0] = ("chr1", 0, 5, b"ACGTN", ReadColumns.empty())
    o, n = run_both(contigs, CallableOptions(min_depth=2, max_depth=4))
    order, fl, sm = o.export()
    assert sm["total_bases"] == sum(c[2] for c in contigs)


@pytest.mark.parametrize("read_len", [12, 36, 70, 71, 151, 250])
def test_synthetic_generator_keeps_every_cigar_op_positive(read_len):
    """synth_short for other read lengths than 150: every op has a positive length, the query lengths match the quality
    strings, reads stay inside the contig (an earlier version wrapped op lengths to 2^28 for reads shorter than 71 bases)."""
    from decodingustools_b200 import synth
    from decodingustools_b200.soa import CONSUMES_QUERY
    c = synth.synth_short("chrS", 40_000, seed=7, depth=20.0, read_len=read_len)
    r = c.reads
    assert r.n > 0 and int((r.cigar >> 4).min()) >= 1 and int((r.cigar >> 4).max()) <= read_len + 30
    qlen = np.add.reduceat(((r.cigar >> 4) * CONSUMES_QUERY[r.cigar & 15]).astype(np.int64), r.cigar_off[:-1].astype(np.int64))
    have = np.diff(r.qual_off.astype(np.int64))
    assert np.all((have == qlen) | (have == 0))                           # SEQ '*' records carry no qualities
    assert int(r.end().max()) <= c.length and np.all(np.diff(r.pos.astype(np.int64)) >= 0)
