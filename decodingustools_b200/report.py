"""Final floating-point aggregation and contig ordering (host side).

Mirrors /root/reference/src/callable_loci/report.rs:15-134 (build_coverage_export), :337-393 (natural
contig order) and profilers/contig_profiler.rs:123-158 (get_quality_stats).  The numbers are plain IEEE
doubles accumulated in the same natural-contig order, so they are bit-identical to the oracle's.
"""
from __future__ import annotations

import functools
from typing import Dict, List

from .callable_loci import CallableProfiler, ContigProfiler
from .options import CalledState


def split_contig_name(name: str):
    for i, ch in enumerate(name):
        if ch.isascii() and (ch.isdigit() or ch in "XYM"):
            return name[:i], name[i:]
    return name, ""


def _order(s: str):
    t = s[1:] if s.startswith("+") else s          # Rust's u32::from_str accepts a leading '+'
    if t and t.isascii() and t.isdigit() and int(t) <= 0xFFFFFFFF:
        return (0, int(t))
    return ({"X": 1, "Y": 2, "M": 3, "MT": 3}.get(s, 4), 0)


def compare_contig_names(a: str, b: str) -> int:
    (ap, asuf), (bp, bsuf) = split_contig_name(a), split_contig_name(b)
    if ap != bp:
        return -1 if ap.encode() < bp.encode() else 1
    (ac, an), (bc, bn) = _order(asuf), _order(bsuf)
    if ac != bc:
        return -1 if ac < bc else 1
    if ac == 0:
        return (an > bn) - (an < bn)
    ab, bb = asuf.encode(), bsuf.encode()
    return (ab > bb) - (ab < bb)


def quality_stats(s: ContigProfiler) -> Dict[str, float]:
    average_mapq = s.summed_mapq / s.quality_bases if s.quality_bases > 0 else 0.0
    average_baseq = s.summed_baseq / s.quality_bases if s.quality_bases > 0 else 0.0
    if s.quality_bases > 0:
        if average_baseq >= 30.0:
            q30 = 100.0
        elif average_baseq < 20.0:
            q30 = 0.0
        else:
            q30 = ((average_baseq - 20.0) / 10.0) * 100.0
    else:
        q30 = 0.0
    return dict(average_mapq=average_mapq, average_baseq=average_baseq, q30_percentage=q30)


def build_coverage_export(contig_stats: Dict[int, ContigProfiler], counter: CallableProfiler, bam_stats: dict | None = None) -> dict:
    """Returns the `export` object of summary.json."""
    bam_stats = bam_stats or {}
    total_bases = callable_bases = q30_bases = total_qpos = total_unique = 0
    total_depth = total_mapq = total_baseq = 0.0
    contigs: List[dict] = []
    ordered = sorted(contig_stats.values(), key=functools.cmp_to_key(lambda a, b: compare_contig_names(a.name, b.name)))
    for s in ordered:
        counts = counter.get_contig_counts(s.name)
        q = quality_stats(s)
        coverage_percent = (s.n_covered_bases / s.length) * 100.0 if s.length > 0 else 0.0
        average_depth = s.summed_coverage / s.n_covered_bases if s.n_covered_bases > 0 else 0.0
        total_bases += s.length
        callable_bases += int(counts[CalledState.CALLABLE])
        total_depth += average_depth * float(s.length)
        total_mapq += q["average_mapq"] * float(s.length)
        total_baseq += q["average_baseq"] * float(s.length)
        q30_bases += int(q["q30_percentage"] / 100.0 * float(s.length))
        total_qpos += s.length
        total_unique += s.n_reads & 0xFFFFFFFF
        contigs.append(dict(
            name=s.name, length=s.length, unique_reads=s.n_reads, coverage_percent=coverage_percent,
            average_depth=average_depth, covered_bases=s.n_covered_bases, total_bases=s.length, quality_stats=q,
            state_distribution=dict(ref_n=int(counts[0]), callable=int(counts[1]), no_coverage=int(counts[2]),
                                    low_coverage=int(counts[3]), excessive_coverage=int(counts[4]),
                                    poor_mapping_quality=int(counts[5]))))
    summary = dict(
        aligner=bam_stats.get("aligner", "Unknown"), reference_build=bam_stats.get("reference_build", "Unknown"),
        sequencing_platform=bam_stats.get("sequencing_platform", "Unknown"), read_length=bam_stats.get("read_length", 0),
        total_bases=total_bases, callable_bases=callable_bases,
        callable_percentage=(callable_bases / total_bases) * 100.0 if total_bases > 0 else 0.0,
        average_depth=total_depth / total_bases if total_bases > 0 else 0.0, contigs_analyzed=len(contig_stats))
    qm = dict(average_mapq=total_mapq / total_qpos if total_qpos > 0 else 0.0,
              average_baseq=total_baseq / total_qpos if total_qpos > 0 else 0.0,
              q30_percentage=(q30_bases / total_qpos) * 100.0 if total_qpos > 0 else 0.0)
    return dict(summary=summary, contigs=contigs, quality_metrics=qm, total_unique_reads=total_unique)


# ------------------------------------------------------------------------------------------------ coverage plot (SVG)
def _xml_attr(v) -> str:
    s = str(v)
    return s.replace("&", "&amp;").replace("'", "&apos;").replace('"', "&quot;").replace("<", "&lt;").replace(">", "&gt;")


def _tag(name: str, attrs, self_closing: bool) -> str:
    """histogram_plotter.rs:34-48.  The reference renders its attributes in HashMap order (different from run to run);
    here they come out in the order the reference code sets them."""
    body = " ".join(f'{k}="{_xml_attr(v)}"' for k, v in attrs)
    return f"<{name} {body}/>" if self_closing else f"<{name} {body}>"


def _u32(v: int) -> int:
    return v & 0xFFFFFFFF


def _bar_height(count: int, stride: int, histogram_height: int) -> int:
    import numpy as np
    h = np.float32(count) / np.float32(stride) * np.float32(histogram_height)      # f32 arithmetic, `as u32` truncates
    return int(h) if h > 0 else 0


def render_coverage_svg(contig_name: str, contig_length: int, stride: int, bins) -> str:
    """The per-contig coverage plot (histogram_plotter.rs:104-410) from the [3][n_bins] bins the device returned
    (rows: CALLABLE, POOR_MAPPING_QUALITY, REF_N; quirk Q2 already added by the BED writer)."""
    callable_d, lowq_d, refn_d = bins[0], bins[1], bins[2]
    svg_width = contig_length // stride
    histogram_height, notch = 100, 10
    header_h = 30 + 15 + 25 + 10
    total_h = header_h + histogram_height + 50
    out = ['<?xml version="1.0" encoding="UTF-8" standalone="no"?>\n']
    out.append(_tag("svg", [("xmlns", "http://www.w3.org/2000/svg"), ("width", svg_width), ("height", total_h),
                            ("style", "background:#%06x" % 0xFFFFFF)], False) + "\n")
    out.append(_tag("text", [("x", svg_width // 2), ("y", 20), ("text-anchor", "middle"), ("font-family", "Arial"), ("font-size", "16"),
                             ("font-weight", "bold"), ("fill", "#000000")], False) + contig_name + "</text>\n")
    out.append(_tag("line", [("x1", 0), ("y1", header_h - 5), ("x2", svg_width), ("y2", header_h - 5), ("stroke", "#808080"),
                             ("stroke-width", 1)], True) + "\n")
    out.append(_tag("rect", [("x", 0), ("y", 0), ("width", svg_width), ("height", header_h - 10), ("fill", "#F8F8F8"), ("opacity", "0.8")], True) + "\n")
    label_y = 30 + 15 + 25 - 5
    for pos in range(0, contig_length + 1, 10_000_000):
        x = pos // stride
        if x >= svg_width:
            continue
        if 20 <= x <= svg_width - 20:
            out.append(_tag("text", [("x", x), ("y", label_y), ("text-anchor", "middle"), ("font-family", "Arial"), ("font-size", "16"),
                                     ("font-weight", "bold"), ("fill", "#800080")], False) + f"{pos // 1_000_000}Mb</text>\n")
        out.append(_tag("line", [("x1", x), ("y1", header_h), ("x2", x), ("y2", header_h + notch), ("stroke", "#800080"), ("stroke-width", 2)], True) + "\n")
        out.append(_tag("line", [("x1", x), ("y1", header_h + histogram_height - notch), ("x2", x), ("y2", header_h + histogram_height),
                                 ("stroke", "#800080"), ("stroke-width", 2)], True) + "\n")
    for x in range(0, contig_length, stride):
        idx = x // stride
        if refn_d[idx] > 0:
            out.append(_tag("rect", [("x", idx), ("y", header_h), ("width", 1), ("height", histogram_height), ("fill", "#000000")], True) + "\n")
            continue
        ch = _bar_height(int(callable_d[idx]), stride, histogram_height) if callable_d[idx] > 0 else 0
        if callable_d[idx] > 0:
            out.append(_tag("rect", [("x", idx), ("y", _u32(header_h + histogram_height - ch)), ("width", 1), ("height", ch), ("fill", "#007700")], True) + "\n")
        if lowq_d[idx] > 0:
            lh = _bar_height(int(lowq_d[idx]), stride, histogram_height)
            out.append(_tag("rect", [("x", idx), ("y", _u32(header_h + histogram_height - lh - ch)), ("width", 1), ("height", lh), ("fill", "#770000")], True) + "\n")
    legend_y = header_h + histogram_height + 10
    lx = _u32(svg_width - 300) // 2                     # wraps for plots narrower than the legend, as the release build does
    out.append("<defs>\n")
    for gid, c0, c1 in (("callableGradient", "#007700", "#00aa00"), ("lowQualGradient", "#770000", "#aa0000")):
        out.append(_tag("linearGradient", [("id", gid), ("x1", "0%"), ("y1", "0%"), ("x2", "100%"), ("y2", "0%"), ("fill", f"url(#{gid})")], False))
        out.append(f'<stop offset="0%" style="stop-color:{c0};stop-opacity:0.8"/>\n')
        out.append(f'<stop offset="100%" style="stop-color:{c1};stop-opacity:0.8"/>\n')
        out.append("</linearGradient>\n")
    out.append("</defs>\n")
    for dx, fill, label in ((0, "url(#callableGradient)", "Callable Coverage"), (150, "url(#lowQualGradient)", "Low Quality Coverage"),
                            (300, "#000000", "Reference N")):
        out.append(_tag("rect", [("x", _u32(lx + dx)), ("y", legend_y), ("width", 20), ("height", 10), ("fill", fill)], True))
        out.append(_tag("text", [("x", _u32(lx + dx + 25)), ("y", legend_y + 8), ("font-family", "Arial"), ("font-size", "12"), ("fill", "#000000")], False))
        out.append(label + "</text>\n")
    out.append("</svg>\n")
    return "".join(out)


# ------------------------------------------------------------------------------------------------ HTML report
# The reference wraps the generated sections in two static template files (templates/report_header.html and
# report_footer.html, include_str!-ed at report.rs:145,154).  They are page furniture, not output of the analysis; this
# package ships its own minimal ones and takes the reference's through `header_html` / `footer_html` when a byte-identical
# page is wanted.
DEFAULT_REPORT_HEADER = """<!DOCTYPE html>
<html lang="en">
<head>
<meta charset="UTF-8">
<title>BAM Analysis Report</title>
<style>
body { font-family: sans-serif; margin: 2rem; }
.stats-columns { display: flex; gap: 3rem; }
dt { font-weight: bold; } dd { margin: 0 0 .5rem 0; }
table { border-collapse: collapse; } td, th { border: 1px solid #ccc; padding: .3rem .6rem; text-align: left; }
.tab-panel { display: none; } .tab-panel.active { display: block; }
.sample-note { font-size: .7em; font-weight: normal; color: #666; }
</style>
</head>
<body>
<main>
<h1>BAM Analysis Report</h1>
"""
DEFAULT_REPORT_FOOTER = """<script>
function switchToContig(id) {
  for (const p of document.querySelectorAll('.tab-panel')) { p.classList.remove('active'); p.style.display = 'none'; }
  const s = document.getElementById(id);
  if (s) { s.classList.add('active'); s.style.display = 'block'; }
}
document.addEventListener('DOMContentLoaded', function () {
  const sel = document.getElementById('contig-select');
  if (sel) switchToContig(sel.value);
});
</script>
</main>
</body>
</html>"""


def _row(label: str, value) -> str:
    return f"<tr><td>{label}</td><td>{value}</td></tr>"


def render_html_report(export: dict, max_samples: int = 10000, header_html: str | None = None, footer_html: str | None = None,
                       plot_exists=None) -> str:
    """report.rs:136-335 (write_html_report): BAM statistics box + one panel per contig in natural order.
    plot_exists(path) decides whether a panel links "<contig>_coverage.svg" (the reference tests the path relative to the
    working directory, report.rs:318-319); default os.path.exists."""
    import os
    plot_exists = plot_exists or os.path.exists
    s, qm = export["summary"], export["quality_metrics"]
    h = [DEFAULT_REPORT_HEADER if header_html is None else header_html]
    h.append("<section class='stats-box'>")
    h.append(f"<h2>BAM Statistics <span class='sample-note'>(based on first {max_samples} reads)</span></h2>")
    h.append("<div class='stats-columns'><dl>")
    pad = "\n            "
    h.append(f"<dt>Reference Build</dt><dd>{s['reference_build']}</dd>{pad}<dt>Aligner</dt><dd>{s['aligner']}</dd>{pad}"
             f"<dt>Sequencing Platform</dt><dd>{s['sequencing_platform']}</dd>{pad}<dt>Average read length</dt><dd>{s['read_length']} bp</dd>{pad}"
             f"<dt>Total Unique Reads</dt><dd>{export['total_unique_reads']}</dd>{pad}<dt>Total Bases</dt><dd>{s['total_bases']}</dd>{pad}")
    h.append("</dl><dl>")
    h.append(f"<dt>Callable Bases</dt><dd>{s['callable_bases']}</dd>{pad}<dt>Callable Percentage</dt><dd>{s['callable_percentage']:.2f}%</dd>{pad}"
             f"<dt>Average Depth</dt><dd>{s['average_depth']:.2f}×</dd>{pad}<dt>Contigs Analyzed</dt><dd>{s['contigs_analyzed']}</dd>{pad}"
             f"<dt>Average MapQ</dt><dd>{qm['average_mapq']:.1f}</dd>{pad}<dt>Average BaseQ</dt><dd>{qm['average_baseq']:.1f}</dd>")
    h.append("</dl></div></section>")
    h.append('<div class="contig-analysis">')
    h.append('<div class="contig-selector">\n        <select id="contig-select" onchange="switchToContig(this.value)" aria-label="Select contig">')
    for i, c in enumerate(export["contigs"]):
        h.append(f'<option value="panel-{i}" {"selected" if i == 0 else ""}>{c["name"]}</option>')
    h.append("</select></div>")
    h.append('<div class="contig-panels">')
    section = '<tr><td colspan="2" style="font-weight: bold; background-color: #f5f5f5;">%s</td></tr>'
    for i, c in enumerate(export["contigs"]):
        q, d = c["quality_stats"], c["state_distribution"]
        h.append(f'<div class="tab-panel {"active" if i == 0 else ""}" id="panel-{i}">')
        h.append("<table><thead><tr><th>Metric</th><th>Value</th></tr></thead><tbody>")
        h.append(_row("Length", f"{c['length']} bp") + _row("Unique Reads", c["unique_reads"]) + _row("Covered Bases", c["covered_bases"]) +
                 _row("Coverage Percent", f"{c['coverage_percent']:.2f}%") + _row("Average Depth", f"{c['average_depth']:.2f}×"))
        h.append(section % "Quality Metrics")
        h.append(_row("Average MapQ", f"{q['average_mapq']:.1f}") + _row("Average BaseQ", f"{q['average_baseq']:.1f}") +
                 _row("Q30 Percentage", f"{q['q30_percentage']:.2f}%"))
        h.append(section % "State Distribution")
        h.append(_row("Reference N", d["ref_n"]) + _row("Callable", d["callable"]) + _row("No Coverage", d["no_coverage"]) +
                 _row("Low Coverage", d["low_coverage"]) + _row("Excessive Coverage", d["excessive_coverage"]) +
                 _row("Poor Mapping Quality", d["poor_mapping_quality"]))
        h.append("</tbody></table>")
        plot = f"{c['name']}_coverage.svg"
        if plot_exists(plot):
            h.append(f'<figure class=\'coverage-plot\'>\n                <img src="{plot}" alt="Coverage distribution for {c["name"]}" loading="lazy">\n'
                     f'                <figcaption>Coverage distribution for {c["name"]}</figcaption>\n            </figure>')
        h.append("</div>")
    h.append("</div></div>")
    h.append(DEFAULT_REPORT_FOOTER if footer_html is None else footer_html)
    return "".join(h)


# ------------------------------------------------------------------------------------------------ summary.json
def format_f64(v: float) -> str:
    """An f64 the way serde_json prints it (ryu's pretty format: shortest round-trip digits, ".0" on integers, plain
    decimals for 1e-5 <= |v| < 1e16, otherwise d.ddde[-]x without a plus sign or padding)."""
    if v == 0:
        return "-0.0" if str(v).startswith("-") else "0.0"
    r = repr(float(v))
    neg = r.startswith("-")
    if neg:
        r = r[1:]
    mant, _, ex = r.partition("e")
    ip, _, fp = mant.partition(".")
    digits = (ip + fp).lstrip("0")
    exp10 = int(ex or 0) + len(ip) - (len(ip + fp) - len((ip + fp).lstrip("0")))    # value = 0.digits * 10^exp10
    digits = digits.rstrip("0") or "0"
    nd, kk = len(digits), exp10                                # kk = position of the decimal point relative to the digits
    if nd <= kk <= 16:
        out = digits + "0" * (kk - nd) + ".0"
    elif 0 < kk <= 16:
        out = digits[:kk] + "." + digits[kk:]
    elif -5 < kk <= 0:
        out = "0." + "0" * (-kk) + digits
    else:
        out = digits[0] + ("." + digits[1:] if nd > 1 else "") + "e" + str(kk - 1)
    return ("-" if neg else "") + out


def _json_str(s: str) -> str:
    return '"' + s.replace("\\", "\\\\").replace('"', '\\"').replace("\n", "\\n").replace("\t", "\\t") + '"'


def render_summary_json(export: dict, bed_file: str, summary_html: str, coverage_plots) -> str:
    """summary.json as main.rs:67-69 writes it (serde_json::to_writer_pretty of CoverageOutput: two-space indent, struct
    field order of export/formats/coverage.rs and api/coverage.rs:134-145, no trailing newline)."""
    s, qm = export["summary"], export["quality_metrics"]

    def contig(c):
        q, d = c["quality_stats"], c["state_distribution"]
        return ("      {\n"
                f'        "name": {_json_str(c["name"])},\n        "length": {c["length"]},\n        "unique_reads": {c["unique_reads"]},\n'
                f'        "coverage_percent": {format_f64(c["coverage_percent"])},\n        "average_depth": {format_f64(c["average_depth"])},\n'
                f'        "covered_bases": {c["covered_bases"]},\n        "total_bases": {c["total_bases"]},\n        "quality_stats": {{\n'
                f'          "average_mapq": {format_f64(q["average_mapq"])},\n          "average_baseq": {format_f64(q["average_baseq"])},\n'
                f'          "q30_percentage": {format_f64(q["q30_percentage"])}\n        }},\n        "state_distribution": {{\n'
                f'          "ref_n": {d["ref_n"]},\n          "callable": {d["callable"]},\n          "no_coverage": {d["no_coverage"]},\n'
                f'          "low_coverage": {d["low_coverage"]},\n          "excessive_coverage": {d["excessive_coverage"]},\n'
                f'          "poor_mapping_quality": {d["poor_mapping_quality"]}\n        }}\n      }}')

    contigs = ",\n".join(contig(c) for c in export["contigs"])
    plots = ",\n".join("      " + _json_str(p) for p in coverage_plots)
    return ("{\n  \"export\": {\n    \"summary\": {\n"
            f'      "aligner": {_json_str(s["aligner"])},\n      "reference_build": {_json_str(s["reference_build"])},\n'
            f'      "sequencing_platform": {_json_str(s["sequencing_platform"])},\n      "read_length": {s["read_length"]},\n'
            f'      "total_bases": {s["total_bases"]},\n      "callable_bases": {s["callable_bases"]},\n'
            f'      "callable_percentage": {format_f64(s["callable_percentage"])},\n      "average_depth": {format_f64(s["average_depth"])},\n'
            f'      "contigs_analyzed": {s["contigs_analyzed"]}\n    }},\n    "contigs": [' + (f"\n{contigs}\n    " if contigs else "") + "],\n"
            f'    "quality_metrics": {{\n      "average_mapq": {format_f64(qm["average_mapq"])},\n      "average_baseq": {format_f64(qm["average_baseq"])},\n'
            f'      "q30_percentage": {format_f64(qm["q30_percentage"])}\n    }},\n    "total_unique_reads": {export["total_unique_reads"]}\n  }},\n'
            f'  "files": {{\n    "bed_file": {_json_str(bed_file)},\n    "summary_html": {_json_str(summary_html)},\n    "coverage_plots": ['
            + (f"\n{plots}\n    " if plots else "") + "]\n  }\n}")
