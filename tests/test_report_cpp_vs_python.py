"""The host-side C++ of the CLI on the CPU, through a g++-built shim (tests/cpp/): the report writer
(csrc/host/report_writer.hpp) against the Python mirror -- platform inference and the BAM sampler on generated read
names, the SVG plot on random bins, the HTML page byte for byte -- and the BGZF/BAM/FASTA reader (csrc/host/bam_reader.hpp)
against files written by tests/bamio.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from decodingustools_b200 import bam_stats as bs
from decodingustools_b200 import report

HERE = os.path.dirname(os.path.abspath(__file__))
PLATFORMS = [bs.ILLUMINA, bs.PACBIO, bs.NANOPORE, bs.MGI, bs.UNKNOWN]        # enum order of report_writer.hpp


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shim") / "report_shim.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", so, os.path.join(HERE, "cpp", "report_shim.cpp"), "-lz", "-lpthread"], check=True)
    L = C.CDLL(so)
    L.shim_detect_platform.restype = C.c_int
    for f in (L.shim_bam_stats, L.shim_svg, L.shim_html):
        f.restype = C.c_size_t
    L.shim_default_header.restype = C.c_char_p; L.shim_default_footer.restype = C.c_char_p
    L.shim_bam_scan.restype = C.c_int; L.shim_load_contig.restype = C.c_int64
    L.shim_fmt_f64.restype = C.c_size_t; L.shim_fmt_f64.argtypes = [C.c_double, C.c_char_p, C.c_size_t]
    L.shim_jstr.restype = C.c_size_t
    L.shim_bam_fetch_tid.restype = C.c_int64
    return L


def _names(rng, n):
    hexd = "0123456789abcdef"
    out = []
    for i in range(n):
        kind = int(rng.integers(0, 9))
        r = lambda a, b: int(rng.integers(a, b))
        h = lambda k: "".join(hexd[r(0, 16)] for _ in range(k))
        if kind == 0:
            out.append(f"{'ADJKENMVFZ'[r(0, 10)]}{r(0, 99999):05d}:{r(1, 500)}:H{h(8).upper()}:{r(1, 5)}:{r(1101, 2678)}:{r(1, 30000)}:{r(1, 30000)}")
        elif kind == 1:
            out.append(f"m{['64', '54', '84', '99'][r(0, 4)]}{r(0, 999):03d}e_{r(200101, 251231)}_{r(0, 235959):06d}/{r(1, 10 ** 6)}/ccs")
        elif kind == 2:
            out.append(f"{h(8)}-{h(4)}-{h(4)}-{h(4)}-{h(12)}" + ("_extra" if r(0, 4) == 0 else ""))
        elif kind == 3:
            out.append(f"{['V300', 'E100', 'CL100', 'G400', 'G99', 'v300'][r(0, 6)]}{r(0, 10 ** 6):06d}L{r(1, 5)}C{r(1, 999):03d}R{r(1, 10 ** 8):08d}")
        elif kind == 4:
            out.append(f"{['V3', 'E1', 'CL1', 'G4', 'X9'][r(0, 5)]}{r(0, 999)}:{r(1, 9)}:{'LX'[r(0, 2)]}{r(1, 99):02d}:{r(1, 5)}:{r(1, 99)}:{r(1, 9999)}:{r(1, 9999)}:{r(0, 2)}")
        elif kind == 5:
            out.append(f"run{r(1, 50)}_ch{r(1, 512)}_read{r(1, 10 ** 5)}_strand{'_pass' * r(0, 3)}")
        elif kind == 6:
            out.append(f"chr{r(1, 23)}:q{r(0, 10 ** 6)}")
        elif kind == 7:
            out.append("".join(":-_/mLCRchread0189AVG"[r(0, 21)] for _ in range(r(1, 48))))
        else:
            out.append(f"{h(8)}-{h(4)}-{h(4)}-{h(4)}-{h(11)}Z")
    return out


def test_platform_detection_and_sampler_agree(shim):
    rng = np.random.default_rng(99)
    names = _names(rng, 4000)
    for q in names:
        assert PLATFORMS[shim.shim_detect_platform(q.encode())] == bs.detect_platform_from_qname(q), q
    for trial in range(40):
        k = int(rng.integers(1, 400))
        # a dominant platform plus some noise, so that the most common instrument is well defined more often than not
        pool = [q for q in names if bs.detect_platform_from_qname(q) == PLATFORMS[trial % 5]] or names
        sub = [pool[int(rng.integers(0, len(pool)))] if rng.random() < 0.8 else names[int(rng.integers(0, len(names)))] for _ in range(k)]
        flags = rng.choice(np.array([0, 1, 16, 0x100, 0x800, 0x900], np.uint16), k)
        lens = rng.integers(0, 20000, k).astype(np.uint64)
        cap = int(rng.integers(1, k + 50))
        py = bs.BamStats(cap).collect(zip(sub, flags.tolist(), lens.tolist()))
        blob = b"".join(q.encode() + b"\0" for q in sub)
        rc, al, pr = C.c_uint64(), C.c_uint64(), C.c_int()
        buf = C.create_string_buffer(256)
        shim.shim_bam_stats(blob, flags.ctypes.data_as(C.c_void_p), lens.ctypes.data_as(C.c_void_p), C.c_uint64(k), C.c_uint64(cap),
                            C.byref(rc), C.byref(al), C.byref(pr), buf, C.c_size_t(256))
        assert (rc.value, al.value, PLATFORMS[pr.value], buf.value.decode()) == (py.read_count, py.average_read_length(), py.primary_platform(), py.infer_platform())


@pytest.mark.parametrize("length,stride,name", [(16_569, 83, "chrM"), (1_000, 10, "t<&>\"'"), (45_000_000, 22_500, "chr9"), (248_956_422, 124_479, "chr1"),
                                                 (5_000, 124_479, "tiny"), (300_001, 1, "wide")])
def test_svg_matches(shim, length, stride, name):
    rng = np.random.default_rng(length % 9973)
    n_bins = length // stride + 1
    bins = np.zeros((3, n_bins), np.uint32)
    for row, p in ((0, 0.7), (1, 0.3), (2, 0.1)):
        m = rng.random(n_bins) < p
        bins[row, m] = rng.integers(1, 2 * stride + 2, int(m.sum()))          # also past a full bar (quirk Q2 can do that)
    expect = report.render_coverage_svg(name, length, stride, bins)
    flat = np.ascontiguousarray(bins)
    need = shim.shim_svg(name.encode(), C.c_uint32(length), C.c_uint32(stride), flat.ctypes.data_as(C.c_void_p), C.c_uint32(n_bins), None, C.c_size_t(0))
    buf = C.create_string_buffer(need + 1)
    shim.shim_svg(name.encode(), C.c_uint32(length), C.c_uint32(stride), flat.ctypes.data_as(C.c_void_p), C.c_uint32(n_bins), buf, C.c_size_t(need + 1))
    assert buf.value.decode() == expect


def test_html_matches(shim):
    rng = np.random.default_rng(5)
    for trial in range(20):
        n = int(rng.integers(0, 6))
        f = lambda: float(rng.choice([0.0, 0.005, 0.125, 29.995, 99.995, 100.0, rng.random() * 100, rng.random() * 1e6]))
        contigs = [dict(name=f"chr{i}", length=int(rng.integers(0, 10 ** 9)), unique_reads=int(rng.integers(0, 10 ** 7)), coverage_percent=f(), average_depth=f(),
                        covered_bases=int(rng.integers(0, 10 ** 9)), total_bases=0, quality_stats=dict(average_mapq=f(), average_baseq=f(), q30_percentage=f()),
                        state_distribution={k: int(rng.integers(0, 10 ** 9)) for k in ("ref_n", "callable", "no_coverage", "low_coverage", "excessive_coverage", "poor_mapping_quality")})
                   for i in range(n)]
        export = dict(summary=dict(aligner="BWA-MEM2", reference_build="GRCh38", sequencing_platform="PacBio Sequel II/IIe", read_length=int(rng.integers(0, 30000)),
                                   total_bases=int(rng.integers(0, 4 * 10 ** 9)), callable_bases=int(rng.integers(0, 4 * 10 ** 9)), callable_percentage=f(),
                                   average_depth=f(), contigs_analyzed=n),
                      contigs=contigs, quality_metrics=dict(average_mapq=f(), average_baseq=f(), q30_percentage=f()), total_unique_reads=int(rng.integers(0, 10 ** 9)))
        plots = {c["name"] for c in contigs if rng.random() < 0.5}
        custom = trial % 2 == 1
        head, foot = ("<h>\n", "</f>") if custom else (None, None)
        expect = report.render_html_report(export, max_samples=10000 + trial, header_html=head, footer_html=foot, plot_exists=lambda p: p[:-len("_coverage.svg")] in plots)
        s = export["summary"]
        su = np.array([s["read_length"], export["total_unique_reads"], s["total_bases"], s["callable_bases"], n, 10000 + trial], np.uint64)
        sd = np.array([s["callable_percentage"], s["average_depth"], export["quality_metrics"]["average_mapq"], export["quality_metrics"]["average_baseq"]], np.float64)
        cu = np.array([[c["length"], c["unique_reads"], c["covered_bases"]] + [c["state_distribution"][k] for k in
                       ("ref_n", "callable", "no_coverage", "low_coverage", "excessive_coverage", "poor_mapping_quality")] + [int(c["name"] in plots)] for c in contigs],
                      np.uint64).reshape(n, 10)
        cd = np.array([[c["coverage_percent"], c["average_depth"], c["quality_stats"]["average_mapq"], c["quality_stats"]["average_baseq"],
                        c["quality_stats"]["q30_percentage"]] for c in contigs], np.float64).reshape(n, 5)
        names = b"".join(c["name"].encode() + b"\0" for c in contigs) or b"\0"
        args = [b"GRCh38", b"BWA-MEM2", b"PacBio Sequel II/IIe", su.ctypes.data_as(C.c_void_p), sd.ctypes.data_as(C.c_void_p), C.c_uint64(n), names,
                cu.ctypes.data_as(C.c_void_p), cd.ctypes.data_as(C.c_void_p), head.encode() if custom else None, foot.encode() if custom else None]
        need = shim.shim_html(*args, None, C.c_size_t(0))
        buf = C.create_string_buffer(need + 1)
        shim.shim_html(*args, buf, C.c_size_t(need + 1))
        assert buf.value.decode() == expect
    assert shim.shim_default_header().decode() == report.DEFAULT_REPORT_HEADER and shim.shim_default_footer().decode() == report.DEFAULT_REPORT_FOOTER


def _scan(shim, path, threads, n_reads, n_cigar, n_qual, n_ref):
    cap_r, cap_c, cap_q = n_reads + 8, n_cigar + 8, n_qual + 8
    tid = np.zeros(cap_r, np.int32); pos = np.zeros(cap_r, np.int32); flag = np.zeros(cap_r, np.uint16); mapq = np.zeros(cap_r, np.uint8)
    co = np.zeros(cap_r + 1, np.uint32); cg = np.zeros(cap_c, np.uint32); qo = np.zeros(cap_r + 1, np.uint64); ql = np.zeros(cap_q, np.uint8)
    names = C.create_string_buffer(64 * cap_r); hdr = C.create_string_buffer(1 << 16); rn = C.create_string_buffer(1 << 12)
    lens = np.zeros(max(n_ref, 1) + 4, np.uint32); nref = C.c_uint32(); n = C.c_uint64(); err = C.create_string_buffer(512)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = shim.shim_bam_scan(path.encode(), C.c_uint(threads), C.c_uint64(cap_r), C.c_uint64(cap_c), C.c_uint64(cap_q), p(tid), p(pos), p(flag), p(mapq),
                            p(co), p(cg), p(qo), p(ql), names, C.c_size_t(len(names)), hdr, C.c_size_t(len(hdr)), rn, C.c_size_t(len(rn)), p(lens),
                            C.byref(nref), C.byref(n), err, C.c_size_t(512))
    if rc != 0:
        raise RuntimeError(err.value.decode())
    k = n.value
    return dict(n=k, tid=tid[:k], pos=pos[:k], flag=flag[:k], mapq=mapq[:k], cigar_off=co[:k + 1], cigar=cg[:int(co[k])], qual_off=qo[:k + 1],
                qual=ql[:int(qo[k])], names=names.value.decode().split("\n")[:k], header=hdr.value.decode(), ref_names=rn.value.decode().split("\n")[:nref.value],
                ref_lens=lens[:nref.value])


@pytest.mark.parametrize("block,threads", [(0xFF00, 4), (777, 3), (97, 1)])
def test_bam_reader_reads_back_what_was_written(shim, tmp_path, block, threads):
    from tests import bamio
    from tests.test_oracle_vs_naive import random_reads
    rng = np.random.default_rng(block)
    contigs = [("chrA", 5000, random_reads(rng, 5000, 300, max_len=80)), ("chrEmpty", 10, random_reads(rng, 10, 0)),
               ("chrB", 900, random_reads(rng, 900, 120, max_len=40))]
    path = str(tmp_path / "t.bam")
    # CIGARs with more than 2 ops go through the CG-tag convention htslib uses past 65535 ops
    bamio.write_bam(path, contigs, block=block, cg_threshold=2, unmapped_tail=5)
    tot = sum(c[2].n for c in contigs)
    got = _scan(shim, path, threads, tot + 5, sum(c[2].n_cigar for c in contigs), sum(c[2].n_qual for c in contigs) + 50, 3)
    assert got["n"] == tot + 5 and got["ref_names"] == ["chrA", "chrEmpty", "chrB"] and got["ref_lens"].tolist() == [5000, 10, 900]
    assert got["header"].startswith("@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:chrA\tLN:5000\n") and "@PG\tID:bwa" in got["header"]
    o = 0
    n_cg = 0
    for tid, (name, _, rc) in enumerate(contigs):
        k = rc.n
        sl = slice(o, o + k)
        assert (got["tid"][sl] == tid).all() and np.array_equal(got["pos"][sl], rc.pos) and np.array_equal(got["flag"][sl], rc.flag) and np.array_equal(got["mapq"][sl], rc.mapq)
        c0, q0 = int(got["cigar_off"][o]), int(got["qual_off"][o])
        assert np.array_equal(got["cigar_off"][o:o + k + 1] - c0, rc.cigar_off) and np.array_equal(got["cigar"][c0:c0 + rc.n_cigar], rc.cigar)   # real CIGARs, not placeholders
        assert np.array_equal(got["qual_off"][o:o + k + 1] - q0, rc.qual_off) and np.array_equal(got["qual"][q0:q0 + rc.n_qual], rc.qual)
        assert got["names"][sl] == [bamio.default_qname(name, int(i)) for i in (rc.name_id if rc.name_id is not None else range(k))]
        n_cg += int((np.diff(rc.cigar_off.astype(np.int64)) > 2).sum())
        o += k
    assert n_cg > 20                                                       # the CG path was exercised
    assert (got["tid"][o:] == -1).all() and (got["flag"][o:] == 4).all() and (got["qual"][-50:] == 0xFF).all() and got["names"][-1] == "unmapped4"


def test_bam_reader_errors_and_fasta(shim, tmp_path):
    from tests import bamio
    bad = tmp_path / "bad.bam"; bad.write_bytes(b"this is not a BGZF file at all, not even close............")
    with pytest.raises(RuntimeError, match="not a BGZF file"):
        _scan(shim, str(bad), 2, 10, 10, 10, 1)
    with pytest.raises(RuntimeError, match="Failed to open BAM file"):
        _scan(shim, str(tmp_path / "missing.bam"), 2, 10, 10, 10, 1)
    notbam = str(tmp_path / "notbam.bam"); bamio.write_bgzf(notbam, b"SAM\1" + bytes(64))
    with pytest.raises(RuntimeError, match="not a BAM"):
        _scan(shim, notbam, 2, 10, 10, 10, 1)
    rng = np.random.default_rng(3)
    seqs = [("chr1", bytes(rng.choice(list(b"ACGTNacgtn"), size=1234).tolist())), ("chrM", b"ACGT" * 17 + b"N"), ("empty", b"")]
    fa = str(tmp_path / "r.fa"); bamio.write_fasta(fa, seqs, width=61)
    for name, seq in seqs:
        buf = np.zeros(len(seq) + 8, np.uint8); err = C.create_string_buffer(256)
        n = shim.shim_load_contig(fa.encode(), name.encode(), buf.ctypes.data_as(C.c_void_p), C.c_uint64(len(buf)), err, C.c_size_t(256))
        assert n == len(seq) and buf[:n].tobytes() == seq
    assert shim.shim_load_contig(fa.encode(), b"chrZ", None, C.c_uint64(0), None, C.c_size_t(0)) == -1


def test_json_scalars_match_and_follow_serde_json(shim):
    """f64 -> text is where summary.json could silently differ from the reference's (serde_json + ryu): known answers for
    the format switches, then C++ against Python on random doubles of every magnitude."""
    known = {0.0: "0.0", 1.0: "1.0", 100.0: "100.0", 0.1: "0.1", 99.5: "99.5", 1e15: "1000000000000000.0", 1e16: "1e16", 1.5e16: "1.5e16",
             1e-5: "0.00001", 1.234e-5: "0.00001234", 9.99e-6: "9.99e-6", 1e-7: "1e-7", 0.30000000000000004: "0.30000000000000004",
             5e-324: "5e-324", 1.7976931348623157e308: "1.7976931348623157e308", 12345.678: "12345.678", -3.25: "-3.25", 41.13333333333333: "41.13333333333333"}
    buf = C.create_string_buffer(64)
    for v, text in known.items():
        shim.shim_fmt_f64(v, buf, 64)
        assert buf.value.decode() == text == report.format_f64(v), v
    rng = np.random.default_rng(11)
    vals = np.concatenate([rng.random(3000) * 10.0 ** rng.integers(-12, 20, 3000), rng.integers(0, 10 ** 9, 500) / 1e3,
                           np.frombuffer(rng.bytes(8 * 3000), dtype=np.float64), rng.integers(0, 2 ** 53, 500).astype(np.float64)])
    for v in vals[np.isfinite(vals)]:
        shim.shim_fmt_f64(float(v), buf, 64)
        t = buf.value.decode()
        assert t == report.format_f64(float(v)) and float(t) == float(v), v
    for text in ["chr1", 'we"ird\\name', "tab\there", "line\nbreak", ""]:
        need = shim.shim_jstr(text.encode(), None, C.c_size_t(0))
        b2 = C.create_string_buffer(need + 1)
        shim.shim_jstr(text.encode(), b2, C.c_size_t(need + 1))
        assert b2.value.decode() == report._json_str(text)


def test_summary_json_layout():
    import json
    from tests.test_report_outputs import _export
    ex = _export()
    text = report.render_summary_json(ex, "out dir/callable.bed", "summary.html", ["chr1_coverage.svg", "chrM_coverage.svg"])
    assert json.loads(text) == dict(export=ex, files=dict(bed_file="out dir/callable.bed", summary_html="summary.html",
                                                         coverage_plots=["chr1_coverage.svg", "chrM_coverage.svg"]))
    assert text.startswith('{\n  "export": {\n    "summary": {\n      "aligner": "BWA",\n') and text.endswith('"chrM_coverage.svg"\n    ]\n  }\n}')
    empty = dict(ex, contigs=[])
    assert '"contigs": [],' in report.render_summary_json(empty, "b", "h", []) and '"coverage_plots": []\n' in report.render_summary_json(empty, "b", "h", [])


@pytest.mark.parametrize("block", [0xFF00, 300])
def test_bai_fetch_jumps_to_each_contig(shim, tmp_path, block):
    from tests import bamio
    from tests.test_oracle_vs_naive import random_reads
    rng = np.random.default_rng(block + 1)
    contigs = [("chrA", 40_000, random_reads(rng, 40_000, 500, max_len=90)), ("chrEmpty", 500, random_reads(rng, 500, 0)),
               ("chrB", 70_000, random_reads(rng, 70_000, 800, max_len=60)), ("chrC", 300, random_reads(rng, 300, 40, max_len=30))]
    path = str(tmp_path / "i.bam")
    bamio.write_bam(path, contigs, block=block, unmapped_tail=3, index=True)
    err = C.create_string_buffer(256)
    for tid, (_, _, rc) in enumerate(contigs):
        ps, fv = C.c_int64(), C.c_uint64()
        n = shim.shim_bam_fetch_tid(path.encode(), C.c_int32(tid), C.byref(ps), C.byref(fv), err, C.c_size_t(256))
        assert n == rc.n, err.value
        if rc.n:
            assert ps.value == int(rc.pos.astype(np.int64).sum())
        else:
            assert fv.value == 2 ** 64 - 1
    os.rename(path + ".bai", str(tmp_path / "i.bai"))                       # <stem>.bai is found as well
    assert shim.shim_bam_fetch_tid(path.encode(), C.c_int32(2), C.byref(ps), C.byref(fv), err, C.c_size_t(256)) == contigs[2][2].n
    os.remove(str(tmp_path / "i.bai"))
    assert shim.shim_bam_fetch_tid(path.encode(), C.c_int32(0), C.byref(ps), C.byref(fv), err, C.c_size_t(256)) == -1
    (tmp_path / "i.bam.bai").write_bytes(b"BAI\1\4\0\0\0\1\0")
    assert shim.shim_bam_fetch_tid(path.encode(), C.c_int32(0), C.byref(ps), C.byref(fv), err, C.c_size_t(256)) == -2 and b"truncated BAI" in err.value


def test_fast_bam_packer_round_trip(tmp_path):
    """clb-pack-bam (bench infrastructure) writes what the Python reader and the C++ reader read back."""
    from decodingustools_b200 import synth
    from tests import bamio
    c = synth.synth_short("chr22", 60_000, seed=77)
    bam = str(tmp_path / "p.bam")
    bamio.pack_bam_fast(bam, c.name, c.length, c.reads, str(tmp_path / "cols"), threads=3)
    text, refs, cols, names = bamio.read_bam(bam)
    assert refs == [("chr22", 60_000)] and "@PG\tID:bwa" in text
    r, w = cols[0], c.reads
    for k in ("pos", "flag", "mapq", "cigar_off", "cigar", "qual_off", "qual"):
        assert np.array_equal(getattr(r, k), getattr(w, k)), k
    assert names[0][0] == f"A00123:7:HFLOWCELLX:1:1101:chr22:{int(w.name_id[0])}"
    assert len(set(names[0])) == len(set(w.name_id.tolist()))
