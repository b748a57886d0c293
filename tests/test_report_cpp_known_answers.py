"""The C++ report writer of the CLI (csrc/host/report_writer.hpp: rows N2-N4) against the HAND-DERIVED known answers of
tests/test_report_outputs.py -- the reference's own doc-comment examples (platform_inference.rs), the worked SVG geometry
(histogram_plotter.rs:104-410) and the HTML fragments (report.rs:162-335) -- directly, not through the Python mirror.
CPU only, through the g++-built shim of tests/cpp/."""
import ctypes as C
import re

import numpy as np

from decodingustools_b200 import bam_stats as bs
from tests.test_report_cpp_vs_python import PLATFORMS, shim  # noqa: F401  (pytest fixture)
from tests.test_report_outputs import _elements, _export


def test_cpp_platform_detection_known_answers(shim):
    cases = [("A00123:123:HXXXYDRXX:1:1101:1000:1000", bs.ILLUMINA), ("m64023e_230414_133043/1/ccs", bs.PACBIO),
             ("0a1b2c3d-4e5f-6a7b-8c9d-0e1f2a3b4c5d", bs.NANOPORE), ("0a1b2c3d-4e5f-6a7b-8c9d-0e1f2a3b4c5g", bs.UNKNOWN),
             ("run1_ch102_read4711_strand_template_pass", bs.NANOPORE), ("V300012345L1C001R00100000001", bs.MGI),
             ("CL100012345L1C001R001_1", bs.MGI), ("E100:1:L01:1:2:3:4:5", bs.MGI), ("E1:1:L1:1:2:3:4", bs.ILLUMINA),
             ("G1234:12:X01:1:2:3:4444", bs.ILLUMINA), ("chr1:q17", bs.UNKNOWN), ("short", bs.UNKNOWN)]
    for qname, platform in cases:
        assert PLATFORMS[shim.shim_detect_platform(qname.encode())] == platform, qname


def _sample(shim, recs, cap):
    blob = b"".join(q.encode() + b"\0" for q, _, _ in recs)
    flags = np.array([f for _, f, _ in recs], np.uint16); lens = np.array([l for _, _, l in recs], np.uint64)
    rc, al, pr = C.c_uint64(), C.c_uint64(), C.c_int()
    buf = C.create_string_buffer(256)
    shim.shim_bam_stats(blob, flags.ctypes.data_as(C.c_void_p), lens.ctypes.data_as(C.c_void_p), C.c_uint64(len(recs)), C.c_uint64(cap),
                        C.byref(rc), C.byref(al), C.byref(pr), buf, C.c_size_t(256))
    return rc.value, al.value, PLATFORMS[pr.value], buf.value.decode()


def test_cpp_bam_sampler_known_answer(shim):
    # five sampled records, two of them not primary; the sixth is past the sample (bam_stats.rs: first max_samples records)
    recs = [("A00123:1:FC1:1:1:1:1", 0x1, 150), ("A00123:1:FC1:1:1:1:2", 0x900, 40), ("m64023e_1_2/1/ccs", 0, 152),
            ("A00123:1:FC1:1:1:1:3", 0x800, 10), ("A00123:1:FC2:1:1:1:4", 0, 149), ("A00123:1:FC1:1:1:1:5", 0, 1000)]
    assert _sample(shim, recs, 5) == (3, (150 + 152 + 149) // 3, bs.ILLUMINA, "NovaSeq")
    assert _sample(shim, [("m84001_1_2/9/ccs", 0, 15000)] * 3, 10000) == (3, 15000, bs.PACBIO, "PacBio Revio")
    assert _sample(shim, [("V300012345L1C001R00100000001", 0, 100)] * 2, 10) == (2, 100, bs.MGI, "MGI DNBSEQ/MGISEQ-2000")


def _svg(shim, name, length, stride, bins):
    flat = np.ascontiguousarray(bins)
    n_bins = bins.shape[1]
    need = shim.shim_svg(name.encode(), C.c_uint32(length), C.c_uint32(stride), flat.ctypes.data_as(C.c_void_p), C.c_uint32(n_bins), None, C.c_size_t(0))
    buf = C.create_string_buffer(need + 1)
    shim.shim_svg(name.encode(), C.c_uint32(length), C.c_uint32(stride), flat.ctypes.data_as(C.c_void_p), C.c_uint32(n_bins), buf, C.c_size_t(need + 1))
    return buf.value.decode()


def test_cpp_coverage_svg_known_answers(shim):
    # 1000 bp contig, stride 10 -> 100 px wide, 101 bins (the worked example of test_report_outputs.py)
    bins = np.zeros((3, 101), np.uint32)
    bins[2, 0] = 10; bins[0, 0] = 5; bins[0, 1] = 10; bins[0, 2] = 3; bins[1, 2] = 4; bins[1, 3] = 1
    svg = _svg(shim, "chrT", 1000, 10, bins)
    assert svg.startswith('<?xml version="1.0" encoding="UTF-8" standalone="no"?>\n<svg ') and svg.endswith("</svg>\n")
    assert _elements(svg, "svg")[0] == {"xmlns": "http://www.w3.org/2000/svg", "width": "100", "height": "230", "style": "background:#ffffff"}
    assert [r for r in _elements(svg, "rect") if r.get("width") == "1"] == [
        {"x": "0", "y": "80", "width": "1", "height": "100", "fill": "#000000"},
        {"x": "1", "y": "80", "width": "1", "height": "100", "fill": "#007700"},
        {"x": "2", "y": "150", "width": "1", "height": "30", "fill": "#007700"},
        {"x": "2", "y": "110", "width": "1", "height": "40", "fill": "#770000"},
        {"x": "3", "y": "170", "width": "1", "height": "10", "fill": "#770000"}]
    assert [l for l in _elements(svg, "line") if l["stroke"] == "#800080"] == [
        {"x1": "0", "y1": "80", "x2": "0", "y2": "90", "stroke": "#800080", "stroke-width": "2"},
        {"x1": "0", "y1": "170", "x2": "0", "y2": "180", "stroke": "#800080", "stroke-width": "2"}]
    assert "Mb</text>" not in svg and ">chrT</text>" in svg and svg.count("<linearGradient ") == 2
    assert [r["x"] for r in _elements(svg, "rect") if r.get("width") == "20"] == [str(((100 - 300) & 0xFFFFFFFF) // 2 + d) for d in (0, 150, 300)]
    # 45 Mbp contig, stride 22500 -> 2000 px: labels at 10..40 Mb, f32 bar heights
    n = 45_000_000 // 22_500 + 1
    bins = np.zeros((3, n), np.uint32)
    bins[0, 7] = 7499; bins[0, 8] = 22_275
    svg = _svg(shim, "chr9", 45_000_000, 22_500, bins)
    assert re.findall(r">(\d+)Mb</text>", svg) == ["10", "20", "30", "40"]
    bars = {r["x"]: r for r in _elements(svg, "rect") if r.get("width") == "1"}
    assert bars["7"]["height"] == "33" and bars["7"]["y"] == "147" and bars["8"]["height"] == "99"
    assert [r["x"] for r in _elements(svg, "rect") if r.get("width") == "20"] == ["850", "1000", "1150"]


def test_cpp_html_report_known_answers(shim):
    export = _export()
    s, contigs = export["summary"], export["contigs"]
    n = len(contigs)
    plots = {"chr1"}
    su = np.array([s["read_length"], export["total_unique_reads"], s["total_bases"], s["callable_bases"], n, 10000], np.uint64)
    sd = np.array([s["callable_percentage"], s["average_depth"], export["quality_metrics"]["average_mapq"], export["quality_metrics"]["average_baseq"]], np.float64)
    cu = np.array([[c["length"], c["unique_reads"], c["covered_bases"]] + [c["state_distribution"][k] for k in
                   ("ref_n", "callable", "no_coverage", "low_coverage", "excessive_coverage", "poor_mapping_quality")] + [int(c["name"] in plots)] for c in contigs],
                  np.uint64).reshape(n, 10)
    cd = np.array([[c["coverage_percent"], c["average_depth"], c["quality_stats"]["average_mapq"], c["quality_stats"]["average_baseq"],
                    c["quality_stats"]["q30_percentage"]] for c in contigs], np.float64).reshape(n, 5)
    names = b"".join(c["name"].encode() + b"\0" for c in contigs)
    args = [s["reference_build"].encode(), s["aligner"].encode(), s["sequencing_platform"].encode(), su.ctypes.data_as(C.c_void_p),
            sd.ctypes.data_as(C.c_void_p), C.c_uint64(n), names, cu.ctypes.data_as(C.c_void_p), cd.ctypes.data_as(C.c_void_p), b"<H>", b"<F>"]
    need = shim.shim_html(*args, None, C.c_size_t(0))
    buf = C.create_string_buffer(need + 1)
    shim.shim_html(*args, buf, C.c_size_t(need + 1))
    html = buf.value.decode()
    assert html.startswith("<H><section class='stats-box'><h2>BAM Statistics <span class='sample-note'>(based on first 10000 reads)</span></h2>")
    assert html.endswith("</div></div><F>")
    assert "<dt>Sequencing Platform</dt><dd>NovaSeq</dd>\n            <dt>Average read length</dt><dd>150 bp</dd>" in html
    assert "<dt>Callable Percentage</dt><dd>41.13%</dd>" in html and "<dt>Average Depth</dt><dd>30.00×</dd>" in html    # 29.995 sits just above the tie in binary
    assert "<dt>Average MapQ</dt><dd>40.0</dd>" in html and "<dt>Average BaseQ</dt><dd>23.5</dd></dl></div></section>" in html
    assert '<option value="panel-0" selected>chr1</option><option value="panel-1" >chrM</option></select></div>' in html
    assert "<tr><td>Length</td><td>2000 bp</td></tr><tr><td>Unique Reads</td><td>77</td></tr><tr><td>Covered Bases</td><td>1990</td></tr>" in html
    assert "<tr><td>Coverage Percent</td><td>99.50%</td></tr><tr><td>Average Depth</td><td>30.12×</td></tr>" in html      # 30.125: exact tie -> even
    assert "<tr><td>Q30 Percentage</td><td>100.00%</td></tr>" in html
    assert "<tr><td>Poor Mapping Quality</td><td>90</td></tr></tbody></table><figure class='coverage-plot'>" in html
    assert html.count("<figure") == 1 and 'alt="Coverage distribution for chr1"' in html
