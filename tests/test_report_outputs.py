"""Host-side report outputs downstream of the device path: BAM sampler + platform
inference, SVG coverage plot, HTML page.  CPU only.  Known answers are worked by hand from the reference's rules
(profilers/platform_inference.rs, utils/histogram_plotter.rs, report.rs) -- its own doc-comment examples where it has them."""
import re

import numpy as np
import pytest

from decodingustools_b200 import bam_stats as bs
from decodingustools_b200 import report


@pytest.mark.parametrize("qname,platform", [
    ("A00123:123:HXXXYDRXX:1:1101:1000:1000", bs.ILLUMINA),                    # platform_inference.rs:86 example
    ("m64023e_230414_133043/1/ccs", bs.PACBIO),                                # :46 example
    ("0a1b2c3d-4e5f-6a7b-8c9d-0e1f2a3b4c5d", bs.NANOPORE),                     # :19 example (UUID)
    ("0a1b2c3d-4e5f-6a7b-8c9d-0e1f2a3b4c5g", bs.UNKNOWN),                      # not hex, no colons
    ("run1_ch102_read4711_strand_template_pass", bs.NANOPORE),                 # > 30 chars with "ch" and "read"
    ("V300012345L1C001R00100000001", bs.MGI),                                  # :55 example
    ("CL100012345L1C001R001_1", bs.MGI),
    ("E100:1:L01:1:2:3:4:5", bs.MGI),                                          # >= 6 colons, E-prefix, third field starts with L
    ("E1:1:L1:1:2:3:4", bs.ILLUMINA),                                          # the MGI rules only look at names longer than 15 chars
    ("G1234:12:X01:1:2:3:4444", bs.ILLUMINA),                                  # G-prefix but the third field does not start with L
    ("chr1:q17", bs.UNKNOWN),
    ("short", bs.UNKNOWN),
])
def test_detect_platform_from_qname(qname, platform):
    assert bs.detect_platform_from_qname(qname) == platform


def test_read_name_parsers():
    assert bs.parse_illumina_read_name("A00123:123:HXXXYDRXX:1:1101:1000:1000") == ("A00123", "HXXXYDRXX")
    assert bs.parse_illumina_read_name("a:b") is None
    assert bs.parse_pacbio_read_name("m64023e_230414_133043/1/ccs") == "m64023e"
    assert bs.parse_pacbio_read_name("x64023e_230414/1") is None
    assert bs.parse_nanopore_read_name("0a1b2c3d-4e5f-6a7b-8c9d-0e1f2a3b4c5d") == "0a1b2c3d"
    assert bs.parse_nanopore_read_name("run7_ch1_read2") == "run7"
    assert bs.parse_nanopore_read_name("plain") == "nanopore"
    assert bs.parse_mgi_read_name("V300012345:L1:C001:R001") == ("V300012345", "L1")
    assert bs.parse_mgi_read_name("V300012345L1C001R00100000001") == ("V300012345", "L1C001")
    assert bs.parse_mgi_read_name("V300012345X1") is None


@pytest.mark.parametrize("primary,instruments,expected", [
    (bs.ILLUMINA, {"A00123": 5, "M0001": 2}, "NovaSeq"), (bs.ILLUMINA, {"M0001": 9, "A1": 2}, "MiSeq"), (bs.ILLUMINA, {"Z9": 1}, "Unknown Illumina"),
    (bs.ILLUMINA, {}, "Unknown Illumina"), (bs.ILLUMINA, {"v1": 1}, "NovaSeq X"),
    (bs.PACBIO, {"m84001": 3}, "PacBio Revio"), (bs.PACBIO, {"m64023e": 3}, "PacBio Sequel II/IIe"), (bs.PACBIO, {"m54001": 3}, "PacBio Sequel"),
    (bs.PACBIO, {"m9": 3}, "PacBio"), (bs.PACBIO, {}, "PacBio"),
    (bs.NANOPORE, {"x": 1}, "Oxford Nanopore"),
    (bs.MGI, {"V300012345": 4}, "MGI DNBSEQ/MGISEQ-2000"), (bs.MGI, {"E100": 4}, "MGI MGISEQ-200"), (bs.MGI, {"CL100x": 4}, "MGI MGISEQ-T7"),
    (bs.MGI, {"G400": 4}, "MGI DNBSEQ-G400"), (bs.MGI, {"G99": 4}, "MGI MGISEQ-T1"), (bs.MGI, {"Q": 4}, "MGI DNBseq"), (bs.MGI, {}, "MGI DNBseq"),
    (bs.UNKNOWN, {"A1": 1}, "Unknown"),
])
def test_infer_specific_platform(primary, instruments, expected):
    assert bs.infer_specific_platform(primary, instruments) == expected


def test_bam_stats_sampler():
    st = bs.BamStats(max_samples=5)
    st.set_header("@HD\tVN:1.6\n@SQ\tSN:chr1\tLN:248956422\n@PG\tID:bwa-mem2\tPN:bwa-mem2\n")
    recs = [("A00123:1:FC1:1:1:1:1", 0x1, 150), ("A00123:1:FC1:1:1:1:2", 0x900, 40), ("m64023e_1_2/1/ccs", 0, 152),
            ("A00123:1:FC1:1:1:1:3", 0x800, 10), ("A00123:1:FC2:1:1:1:4", 0, 149), ("A00123:1:FC1:1:1:1:5", 0, 1000)]
    st.collect(recs)
    assert st.seen == 5 and st.read_count == 3                       # the sixth record is past the sample, two are not primary
    assert st.average_read_length() == (150 + 152 + 149) // 3
    assert st.platform_counts == {bs.ILLUMINA: 2, bs.PACBIO: 1} and st.instruments == {"A00123": 2, "m64023e": 1}
    assert st.flow_cells == {"FC1": 1, "FC2": 1} and st.paired_reads == 1
    assert st.infer_platform() == "NovaSeq"
    assert st.as_summary_fields() == dict(aligner="BWA-MEM2", reference_build="GRCh38", sequencing_platform="NovaSeq", read_length=150)
    assert bs.BamStats().infer_platform() == "Unknown" and bs.BamStats().average_read_length() == 0


def _elements(svg, name):
    return [dict(re.findall(r'([\w-]+)="([^"]*)"', m)) for m in re.findall(rf"<{name} ([^>]*?)/?>", svg)]


def test_coverage_svg_known_answer():
    # 1000 bp contig, stride 10 -> 100 px wide, 101 bins
    bins = np.zeros((3, 101), np.uint32)
    bins[2, 0] = 10                      # REF_N column: full-height black bar, nothing else drawn there
    bins[0, 0] = 5
    bins[0, 1] = 10                      # all callable: 100 px
    bins[0, 2] = 3; bins[1, 2] = 4       # 30 px green with 40 px red stacked on top
    bins[1, 3] = 1                       # 10 px red from the floor
    svg = report.render_coverage_svg("chrT", 1000, 10, bins)
    assert svg.startswith('<?xml version="1.0" encoding="UTF-8" standalone="no"?>\n<svg ') and svg.endswith("</svg>\n")
    root = _elements(svg, "svg")[0]
    assert root == {"xmlns": "http://www.w3.org/2000/svg", "width": "100", "height": "230", "style": "background:#ffffff"}
    bars = [r for r in _elements(svg, "rect") if r.get("width") == "1"]
    assert bars == [
        {"x": "0", "y": "80", "width": "1", "height": "100", "fill": "#000000"},
        {"x": "1", "y": "80", "width": "1", "height": "100", "fill": "#007700"},
        {"x": "2", "y": "150", "width": "1", "height": "30", "fill": "#007700"},
        {"x": "2", "y": "110", "width": "1", "height": "40", "fill": "#770000"},
        {"x": "3", "y": "170", "width": "1", "height": "10", "fill": "#770000"},
    ]
    # one position mark at 0 (no label: closer than 20 px to the edge), legend x wraps below 300 px like the release build
    assert [l for l in _elements(svg, "line") if l["stroke"] == "#800080"] == [
        {"x1": "0", "y1": "80", "x2": "0", "y2": "90", "stroke": "#800080", "stroke-width": "2"},
        {"x1": "0", "y1": "170", "x2": "0", "y2": "180", "stroke": "#800080", "stroke-width": "2"}]
    assert "Mb</text>" not in svg
    legend = [r for r in _elements(svg, "rect") if r.get("width") == "20"]
    assert [r["x"] for r in legend] == [str(((100 - 300) & 0xFFFFFFFF) // 2 + d) for d in (0, 150, 300)]
    assert ">chrT</text>" in svg and svg.count("<linearGradient ") == 2


def test_coverage_svg_labels_and_f32_heights():
    # 45 Mbp contig, stride 22500 -> 2000 px: marks at 0, 10, 20, 30, 40 Mb; labels where 20 <= x <= width - 20
    n = 45_000_000 // 22_500 + 1
    bins = np.zeros((3, n), np.uint32)
    bins[0, 7] = 7499                    # 7499 / 22500 * 100 = 33.3288.. -> 33
    bins[0, 8] = 22_275                  # 99.0 exactly in f32 arithmetic -> 99
    svg = report.render_coverage_svg("chr9", 45_000_000, 22_500, bins)
    assert re.findall(r">(\d+)Mb</text>", svg) == ["10", "20", "30", "40"]
    bars = {r["x"]: r for r in _elements(svg, "rect") if r.get("width") == "1"}
    assert bars["7"]["height"] == "33" and bars["7"]["y"] == "147" and bars["8"]["height"] == "99"
    lg = [r for r in _elements(svg, "rect") if r.get("width") == "20"]
    assert [r["x"] for r in lg] == ["850", "1000", "1150"]


def _export():
    return dict(
        summary=dict(aligner="BWA", reference_build="GRCh38", sequencing_platform="NovaSeq", read_length=150, total_bases=3000, callable_bases=1234,
                     callable_percentage=41.13333333333333, average_depth=29.995, contigs_analyzed=2),
        contigs=[dict(name="chr1", length=2000, unique_reads=77, coverage_percent=99.5, average_depth=30.125, covered_bases=1990, total_bases=2000,
                      quality_stats=dict(average_mapq=59.96, average_baseq=35.25, q30_percentage=100.0),
                      state_distribution=dict(ref_n=10, callable=1200, no_coverage=0, low_coverage=700, excessive_coverage=0, poor_mapping_quality=90)),
                 dict(name="chrM", length=1000, unique_reads=5, coverage_percent=0.0, average_depth=0.0, covered_bases=0, total_bases=1000,
                      quality_stats=dict(average_mapq=0.0, average_baseq=0.0, q30_percentage=0.0),
                      state_distribution=dict(ref_n=0, callable=34, no_coverage=966, low_coverage=0, excessive_coverage=0, poor_mapping_quality=0))],
        quality_metrics=dict(average_mapq=39.97, average_baseq=23.5, q30_percentage=66.7), total_unique_reads=82)


def test_html_report_sections():
    html = report.render_html_report(_export(), header_html="<H>", footer_html="<F>", plot_exists=lambda p: p == "chr1_coverage.svg")
    assert html.startswith("<H><section class='stats-box'><h2>BAM Statistics <span class='sample-note'>(based on first 10000 reads)</span></h2>")
    assert html.endswith("</div></div><F>")
    assert "<dt>Sequencing Platform</dt><dd>NovaSeq</dd>\n            <dt>Average read length</dt><dd>150 bp</dd>" in html
    assert "<dt>Callable Percentage</dt><dd>41.13%</dd>" in html and "<dt>Average Depth</dt><dd>30.00×</dd>" in html    # 29.995 sits just above the tie in binary
    assert "<dt>Average MapQ</dt><dd>40.0</dd>" in html and "<dt>Average BaseQ</dt><dd>23.5</dd></dl></div></section>" in html
    assert '<option value="panel-0" selected>chr1</option><option value="panel-1" >chrM</option></select></div>' in html
    assert '<div class="tab-panel active" id="panel-0"><table>' in html and '<div class="tab-panel " id="panel-1"><table>' in html
    assert "<tr><td>Length</td><td>2000 bp</td></tr><tr><td>Unique Reads</td><td>77</td></tr><tr><td>Covered Bases</td><td>1990</td></tr>" in html
    assert "<tr><td>Coverage Percent</td><td>99.50%</td></tr><tr><td>Average Depth</td><td>30.12×</td></tr>" in html      # 30.125: exact tie -> even
    assert "<tr><td>Q30 Percentage</td><td>100.00%</td></tr>" in html
    assert "<tr><td>Poor Mapping Quality</td><td>90</td></tr></tbody></table><figure class='coverage-plot'>" in html
    assert html.count("<figure") == 1 and 'alt="Coverage distribution for chr1"' in html
    # the built-in page furniture carries what the generated sections need
    page = report.render_html_report(_export(), plot_exists=lambda p: False)
    assert page.startswith("<!DOCTYPE html>") and "function switchToContig" in page and page.endswith("</html>") and "<figure" not in page
