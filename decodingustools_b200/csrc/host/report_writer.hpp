// report_writer.hpp -- host-side text outputs of the `coverage` command that sit downstream of the device path:
// the BAM sampler + platform inference that fill the report header, the per-contig SVG coverage plot and the HTML page.
//
// Reference behaviour restated here (nothing is executed from it):
//   BamStats sampler                        src/callable_loci/profilers/bam_stats.rs:44-141,185-241
//   platform / instrument inference         src/callable_loci/profilers/platform_inference.rs:16-294
//   SVG plot                                src/callable_loci/utils/histogram_plotter.rs:104-410,412-454
//   HTML report                             src/callable_loci/report.rs:136-335
// decodingustools_b200/bam_stats.py and report.py hold the same logic for Python callers; tests/test_cli_gpu.py compares
// the two byte for byte.
#pragma once
#include <charconv>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

namespace report {

// ------------------------------------------------------------------------------------------------ platform inference
enum Platform { ILLUMINA = 0, PACBIO, NANOPORE, MGI, UNKNOWN, N_PLATFORMS };

inline bool starts_with(const std::string &s, const char *p) { return s.compare(0, strlen(p), p) == 0; }
inline size_t count_char(const std::string &s, char c) { size_t n = 0; for (char x : s) n += x == c; return n; }
inline std::vector<std::string> split(const std::string &s, char sep) {
    std::vector<std::string> out; size_t a = 0;
    for (;;) { const size_t b = s.find(sep, a); if (b == std::string::npos) { out.push_back(s.substr(a)); break; } out.push_back(s.substr(a, b - a)); a = b + 1; }
    return out;
}
inline bool all_hex(const std::string &s) { for (unsigned char c : s) if (!isxdigit(c)) return false; return true; }

inline Platform detect_platform_from_qname(const std::string &q) {                   // platform_inference.rs:16-94
    const bool dash = q.find('-') != std::string::npos, us = q.find('_') != std::string::npos;
    if (q.size() > 30 && (dash || us)) {
        const auto parts = split(q, '-');
        if (parts.size() == 5 && parts[0].size() == 8 && parts[1].size() == 4 && parts[2].size() == 4 && parts[3].size() == 4 && parts[4].size() >= 12) {
            bool hex = true; for (const auto &p : parts) hex = hex && all_hex(p);
            if (hex) return NANOPORE;
        }
        if (q.find("ch") != std::string::npos && q.find("read") != std::string::npos) return NANOPORE;
    }
    if (!q.empty() && q[0] == 'm' && q.find('/') != std::string::npos) {
        const auto parts = split(q, '/');
        if (parts.size() >= 2 && parts[0].find('_') != std::string::npos) return PACBIO;
    }
    if (q.size() > 15) {
        std::string prefix = q.substr(0, 5); for (auto &c : prefix) c = (char)toupper((unsigned char)c);
        if (starts_with(prefix, "V300") || starts_with(prefix, "E100") || starts_with(prefix, "CL100") || starts_with(prefix, "G400") || starts_with(prefix, "G99"))
            return MGI;
        if (count_char(q, ':') >= 6) {
            const auto parts = split(q, ':');
            if ((starts_with(parts[0], "V") || starts_with(parts[0], "E") || starts_with(parts[0], "CL") || starts_with(parts[0], "G")) &&
                parts.size() >= 3 && starts_with(parts[2], "L"))
                return MGI;
        }
    }
    if (count_char(q, ':') >= 6) return ILLUMINA;
    return UNKNOWN;
}

// insertion-ordered counter: the reference's HashMap leaves a tie in max_by_key to hash order; here the first key seen wins
struct Counter {
    std::vector<std::pair<std::string, uint64_t>> items;
    void add(const std::string &k) { for (auto &it : items) if (it.first == k) { it.second++; return; } items.emplace_back(k, 1); }
    const std::string *most_common() const {
        const std::pair<std::string, uint64_t> *best = nullptr;
        for (const auto &it : items) if (!best || it.second > best->second) best = &it;
        return best ? &best->first : nullptr;
    }
};

inline std::string infer_specific_platform(Platform primary, const Counter &instruments) {    // platform_inference.rs:217-293
    const std::string *top = instruments.most_common();
    switch (primary) {
    case PACBIO:
        if (!top) return "PacBio";
        return starts_with(*top, "m84") ? "PacBio Revio" : starts_with(*top, "m64") ? "PacBio Sequel II/IIe" : starts_with(*top, "m54") ? "PacBio Sequel" : "PacBio";
    case NANOPORE: return "Oxford Nanopore";
    case MGI:
        if (!top) return "MGI DNBseq";
        return starts_with(*top, "V300") ? "MGI DNBSEQ/MGISEQ-2000" : starts_with(*top, "E100") ? "MGI MGISEQ-200" : starts_with(*top, "CL100") ? "MGI MGISEQ-T7"
             : starts_with(*top, "G400") ? "MGI DNBSEQ-G400" : starts_with(*top, "G99") ? "MGI MGISEQ-T1" : "MGI DNBseq";
    case ILLUMINA: {
        if (!top || top->empty()) return "Unknown Illumina";
        switch (tolower((unsigned char)(*top)[0])) {
        case 'a': return "NovaSeq";    case 'd': return "HiSeq 2500"; case 'j': return "HiSeq 3000"; case 'k': return "HiSeq 4000";
        case 'e': return "HiSeq X";    case 'n': return "NextSeq";    case 'm': return "MiSeq";      case 'v': return "NovaSeq X";
        case 'f': return "iSeq";       default: return "Unknown Illumina";
        }
    }
    default: return "Unknown";
    }
}

struct BamStats {                                                                     // bam_stats.rs:9-141
    uint64_t max_samples = 10000, seen = 0, read_count = 0, total_read_length = 0;
    Counter instruments, flow_cells;
    uint64_t platform_counts[N_PLATFORMS] = {0, 0, 0, 0, 0};
    int platform_first_seen[N_PLATFORMS] = {-1, -1, -1, -1, -1}; int n_platforms_seen = 0;

    bool full() const { return seen >= max_samples; }
    void add_record(const std::string &q, uint16_t flag, uint64_t seq_len) {
        if (full()) return;
        seen++;
        if (flag & 0x900) return;                                                     // primary alignments only
        read_count++; total_read_length += seq_len;
        const Platform p = detect_platform_from_qname(q);
        if (platform_counts[p]++ == 0) platform_first_seen[p] = n_platforms_seen++;
        if (p == ILLUMINA) {                                                          // platform_inference.rs:99-108
            const auto parts = split(q, ':');
            if (parts.size() >= 3) { instruments.add(parts[0]); flow_cells.add(parts[2]); }
        } else if (p == PACBIO) {                                                     // :114-126
            const std::string movie = q.substr(0, q.find('/'));
            const size_t us = movie.find('_');
            if (!movie.empty() && movie[0] == 'm' && us != std::string::npos) instruments.add(movie.substr(0, us));
        } else if (p == NANOPORE) {                                                   // :132-159
            if (q.size() > 30 && q.find('-') != std::string::npos && split(q, '-').size() >= 5) instruments.add(split(split(q, '_')[0], '-')[0]);
            else if (q.find('_') != std::string::npos) instruments.add(q.substr(0, q.find('_')));
            else instruments.add("nanopore");
        } else if (p == MGI) {                                                        // :165-191
            if (count_char(q, ':') >= 3) { const auto parts = split(q, ':'); instruments.add(parts[0]); flow_cells.add(parts[1]); }
            else if (q.size() > 10) {
                const size_t l = q.find('L');
                if (l != std::string::npos) {
                    const std::string rest = q.substr(l);
                    if (rest.find('C') != std::string::npos) {
                        const size_t r = rest.find('R');
                        instruments.add(q.substr(0, l)); flow_cells.add(rest.substr(0, r == std::string::npos ? rest.size() : r));
                    }
                }
            }
        }
    }
    uint64_t average_read_length() const { return read_count ? total_read_length / read_count : 0; }
    Platform primary_platform() const {
        int best = -1;
        for (int p = 0; p < N_PLATFORMS; p++) {
            if (!platform_counts[p]) continue;
            if (best < 0 || platform_counts[p] > platform_counts[best] ||
                (platform_counts[p] == platform_counts[best] && platform_first_seen[p] < platform_first_seen[best])) best = p;
        }
        return best < 0 ? UNKNOWN : (Platform)best;
    }
    std::string infer_platform() const { return infer_specific_platform(primary_platform(), instruments); }
};

// ------------------------------------------------------------------------------------------------ SVG coverage plot
inline std::string xml_attr(const std::string &s) {
    std::string o;
    for (char c : s) { if (c == '&') o += "&amp;"; else if (c == '\'') o += "&apos;"; else if (c == '"') o += "&quot;"; else if (c == '<') o += "&lt;"; else if (c == '>') o += "&gt;"; else o += c; }
    return o;
}
struct Tag {                       // attributes in the order the reference code sets them (its own output order is HashMap order)
    std::string s;
    explicit Tag(const char *name) : s(std::string("<") + name) {}
    Tag &a(const char *k, const std::string &v) { s += " " + std::string(k) + "=\"" + xml_attr(v) + "\""; return *this; }
    Tag &a(const char *k, uint64_t v) { return a(k, std::to_string(v)); }
    std::string open() const { return s + ">"; }
    std::string closed() const { return s + "/>"; }
};
inline uint32_t bar_height(uint32_t count, uint32_t stride, uint32_t histogram_height) {
    const volatile float q = (float)count / (float)stride;                            // f32 arithmetic, `as u32` truncates
    const float h = q * (float)histogram_height;
    return h > 0.0f ? (uint32_t)h : 0u;
}

// bins: [3][n_bins] rows CALLABLE, POOR_MAPPING_QUALITY, REF_N (quirk Q2 included), n_bins = contig_length / stride + 1
inline std::string render_coverage_svg(const std::string &contig, uint32_t contig_length, uint32_t stride, const uint32_t *bins, uint32_t n_bins) {
    const uint32_t *callable_d = bins, *lowq_d = bins + n_bins, *refn_d = bins + 2 * (size_t)n_bins;
    const uint32_t svg_width = contig_length / stride, hist_h = 100, notch = 10, header_h = 30 + 15 + 25 + 10, total_h = header_h + hist_h + 50;
    std::string o = "<?xml version=\"1.0\" encoding=\"UTF-8\" standalone=\"no\"?>\n";
    o += Tag("svg").a("xmlns", "http://www.w3.org/2000/svg").a("width", svg_width).a("height", total_h).a("style", "background:#ffffff").open() + "\n";
    o += Tag("text").a("x", svg_width / 2).a("y", 20).a("text-anchor", "middle").a("font-family", "Arial").a("font-size", "16").a("font-weight", "bold")
             .a("fill", "#000000").open() + contig + "</text>\n";
    o += Tag("line").a("x1", 0).a("y1", header_h - 5).a("x2", svg_width).a("y2", header_h - 5).a("stroke", "#808080").a("stroke-width", 1).closed() + "\n";
    o += Tag("rect").a("x", 0).a("y", 0).a("width", svg_width).a("height", header_h - 10).a("fill", "#F8F8F8").a("opacity", "0.8").closed() + "\n";
    const uint32_t label_y = 30 + 15 + 25 - 5;
    for (uint64_t pos = 0; pos <= contig_length; pos += 10000000ull) {
        const int64_t x = (int64_t)(pos / stride);
        if (x >= (int64_t)svg_width) continue;
        if (x >= 20 && x <= (int64_t)svg_width - 20)
            o += Tag("text").a("x", (uint64_t)x).a("y", label_y).a("text-anchor", "middle").a("font-family", "Arial").a("font-size", "16").a("font-weight", "bold")
                     .a("fill", "#800080").open() + std::to_string(pos / 1000000ull) + "Mb</text>\n";
        o += Tag("line").a("x1", (uint64_t)x).a("y1", header_h).a("x2", (uint64_t)x).a("y2", header_h + notch).a("stroke", "#800080").a("stroke-width", 2).closed() + "\n";
        o += Tag("line").a("x1", (uint64_t)x).a("y1", header_h + hist_h - notch).a("x2", (uint64_t)x).a("y2", header_h + hist_h).a("stroke", "#800080")
                 .a("stroke-width", 2).closed() + "\n";
    }
    for (uint64_t x = 0; x < contig_length; x += stride) {
        const uint32_t idx = (uint32_t)(x / stride);
        if (refn_d[idx] > 0) { o += Tag("rect").a("x", idx).a("y", header_h).a("width", 1).a("height", hist_h).a("fill", "#000000").closed() + "\n"; continue; }
        const uint32_t ch = callable_d[idx] > 0 ? bar_height(callable_d[idx], stride, hist_h) : 0u;
        if (callable_d[idx] > 0) o += Tag("rect").a("x", idx).a("y", (uint32_t)(header_h + hist_h - ch)).a("width", 1).a("height", ch).a("fill", "#007700").closed() + "\n";
        if (lowq_d[idx] > 0) {
            const uint32_t lh = bar_height(lowq_d[idx], stride, hist_h);
            o += Tag("rect").a("x", idx).a("y", (uint32_t)(header_h + hist_h - lh - ch)).a("width", 1).a("height", lh).a("fill", "#770000").closed() + "\n";
        }
    }
    const uint32_t legend_y = header_h + hist_h + 10;
    const uint32_t lx = (uint32_t)(svg_width - 300u) / 2u;              // wraps for plots narrower than the legend, as the release build does
    o += "<defs>\n";
    const char *grads[2][3] = {{"callableGradient", "#007700", "#00aa00"}, {"lowQualGradient", "#770000", "#aa0000"}};
    for (auto &g : grads) {
        o += Tag("linearGradient").a("id", g[0]).a("x1", "0%").a("y1", "0%").a("x2", "100%").a("y2", "0%").a("fill", std::string("url(#") + g[0] + ")").open();
        o += std::string("<stop offset=\"0%\" style=\"stop-color:") + g[1] + ";stop-opacity:0.8\"/>\n";
        o += std::string("<stop offset=\"100%\" style=\"stop-color:") + g[2] + ";stop-opacity:0.8\"/>\n";
        o += "</linearGradient>\n";
    }
    o += "</defs>\n";
    const struct { uint32_t dx; const char *fill, *label; } legend[3] = {{0, "url(#callableGradient)", "Callable Coverage"},
                                                                         {150, "url(#lowQualGradient)", "Low Quality Coverage"}, {300, "#000000", "Reference N"}};
    for (auto &l : legend) {
        o += Tag("rect").a("x", (uint32_t)(lx + l.dx)).a("y", legend_y).a("width", 20).a("height", 10).a("fill", l.fill).closed();
        o += Tag("text").a("x", (uint32_t)(lx + l.dx + 25)).a("y", legend_y + 8).a("font-family", "Arial").a("font-size", "12").a("fill", "#000000").open();
        o += std::string(l.label) + "</text>\n";
    }
    o += "</svg>\n";
    return o;
}

// ------------------------------------------------------------------------------------------------ summary.json scalars
// serde_json / ryu formatting of an f64: shortest round-trip digits, ".0" for integers, exponent outside 1e-5..1e16
inline std::string fmt_f64(double v) {
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);
    std::string s(buf, res.ptr);                           // d[.ddd]e[+-]xx
    const size_t epos = s.find('e');
    std::string mant = s.substr(0, epos); const int exp10 = atoi(s.c_str() + epos + 1);
    const bool neg = mant[0] == '-'; if (neg) mant.erase(0, 1);
    std::string digits; for (char ch : mant) if (ch != '.') digits.push_back(ch);
    const int nd = (int)digits.size(), kk = exp10 + 1;     // decimal point position relative to digits
    std::string out;
    if (v == 0) out = "0.0";
    else if (nd <= kk && kk <= 16) { out = digits + std::string((size_t)(kk - nd), '0') + ".0"; }
    else if (0 < kk && kk <= 16) { out = digits.substr(0, (size_t)kk) + "." + digits.substr((size_t)kk); }
    else if (-5 < kk && kk <= 0) { out = "0." + std::string((size_t)(-kk), '0') + digits; }
    else { out = digits.substr(0, 1) + (nd > 1 ? "." + digits.substr(1) : "") + "e" + std::to_string(exp10); }
    return (neg ? "-" : "") + out;
}
inline std::string jstr(const std::string &s) {
    std::string o = "\"";
    for (char ch : s) { if (ch == '"' || ch == '\\') { o += '\\'; o += ch; } else if (ch == '\n') o += "\\n"; else if (ch == '\t') o += "\\t"; else o += ch; }
    return o + "\"";
}

// ------------------------------------------------------------------------------------------------ HTML report
// Page furniture: the reference include_str!s two static template files (report.rs:145,154).  These are this tool's own;
// `--report-templates DIR` makes the CLI read report_header.html / report_footer.html from a reference checkout instead.
inline const char *default_header() {
    return "<!DOCTYPE html>\n<html lang=\"en\">\n<head>\n<meta charset=\"UTF-8\">\n<title>BAM Analysis Report</title>\n<style>\n"
           "body { font-family: sans-serif; margin: 2rem; }\n.stats-columns { display: flex; gap: 3rem; }\n"
           "dt { font-weight: bold; } dd { margin: 0 0 .5rem 0; }\n"
           "table { border-collapse: collapse; } td, th { border: 1px solid #ccc; padding: .3rem .6rem; text-align: left; }\n"
           ".tab-panel { display: none; } .tab-panel.active { display: block; }\n"
           ".sample-note { font-size: .7em; font-weight: normal; color: #666; }\n</style>\n</head>\n<body>\n<main>\n<h1>BAM Analysis Report</h1>\n";
}
inline const char *default_footer() {
    return "<script>\nfunction switchToContig(id) {\n"
           "  for (const p of document.querySelectorAll('.tab-panel')) { p.classList.remove('active'); p.style.display = 'none'; }\n"
           "  const s = document.getElementById(id);\n  if (s) { s.classList.add('active'); s.style.display = 'block'; }\n}\n"
           "document.addEventListener('DOMContentLoaded', function () {\n  const sel = document.getElementById('contig-select');\n"
           "  if (sel) switchToContig(sel.value);\n});\n</script>\n</main>\n</body>\n</html>";
}

struct ContigRow {
    std::string name; uint64_t length, unique_reads, covered_bases; double coverage_percent, average_depth, average_mapq, average_baseq, q30_percentage;
    uint64_t counts[6]; bool has_plot;
};
struct Summary {
    std::string reference_build, aligner, sequencing_platform; uint64_t read_length, total_unique_reads, total_bases, callable_bases, contigs_analyzed, max_samples;
    double callable_percentage, average_depth, average_mapq, average_baseq;
};
inline std::string fixed(double v, int prec) { char b[400]; snprintf(b, sizeof b, "%.*f", prec, v); return b; }
inline std::string row(const char *label, const std::string &v) { return std::string("<tr><td>") + label + "</td><td>" + v + "</td></tr>"; }
inline std::string row(const char *label, uint64_t v) { return row(label, std::to_string(v)); }

inline std::string render_html_report(const Summary &s, const std::vector<ContigRow> &contigs, const std::string &header, const std::string &footer) {
    const std::string pad = "\n            ";
    std::string h = header;
    h += "<section class='stats-box'>";
    h += "<h2>BAM Statistics <span class='sample-note'>(based on first " + std::to_string(s.max_samples) + " reads)</span></h2>";
    h += "<div class='stats-columns'><dl>";
    h += "<dt>Reference Build</dt><dd>" + s.reference_build + "</dd>" + pad + "<dt>Aligner</dt><dd>" + s.aligner + "</dd>" + pad + "<dt>Sequencing Platform</dt><dd>" +
         s.sequencing_platform + "</dd>" + pad + "<dt>Average read length</dt><dd>" + std::to_string(s.read_length) + " bp</dd>" + pad + "<dt>Total Unique Reads</dt><dd>" +
         std::to_string(s.total_unique_reads) + "</dd>" + pad + "<dt>Total Bases</dt><dd>" + std::to_string(s.total_bases) + "</dd>" + pad;
    h += "</dl><dl>";
    h += "<dt>Callable Bases</dt><dd>" + std::to_string(s.callable_bases) + "</dd>" + pad + "<dt>Callable Percentage</dt><dd>" + fixed(s.callable_percentage, 2) + "%</dd>" + pad +
         "<dt>Average Depth</dt><dd>" + fixed(s.average_depth, 2) + "\xC3\x97</dd>" + pad + "<dt>Contigs Analyzed</dt><dd>" + std::to_string(s.contigs_analyzed) + "</dd>" + pad +
         "<dt>Average MapQ</dt><dd>" + fixed(s.average_mapq, 1) + "</dd>" + pad + "<dt>Average BaseQ</dt><dd>" + fixed(s.average_baseq, 1) + "</dd>";
    h += "</dl></div></section>";
    h += "<div class=\"contig-analysis\">";
    h += "<div class=\"contig-selector\">\n        <select id=\"contig-select\" onchange=\"switchToContig(this.value)\" aria-label=\"Select contig\">";
    for (size_t i = 0; i < contigs.size(); i++)
        h += "<option value=\"panel-" + std::to_string(i) + "\" " + (i == 0 ? "selected" : "") + ">" + contigs[i].name + "</option>";
    h += "</select></div>";
    h += "<div class=\"contig-panels\">";
    const std::string sec0 = "<tr><td colspan=\"2\" style=\"font-weight: bold; background-color: #f5f5f5;\">", sec1 = "</td></tr>";
    for (size_t i = 0; i < contigs.size(); i++) {
        const ContigRow &c = contigs[i];
        h += std::string("<div class=\"tab-panel ") + (i == 0 ? "active" : "") + "\" id=\"panel-" + std::to_string(i) + "\">";
        h += "<table><thead><tr><th>Metric</th><th>Value</th></tr></thead><tbody>";
        h += row("Length", std::to_string(c.length) + " bp") + row("Unique Reads", c.unique_reads) + row("Covered Bases", c.covered_bases) +
             row("Coverage Percent", fixed(c.coverage_percent, 2) + "%") + row("Average Depth", fixed(c.average_depth, 2) + "\xC3\x97");
        h += sec0 + "Quality Metrics" + sec1;
        h += row("Average MapQ", fixed(c.average_mapq, 1)) + row("Average BaseQ", fixed(c.average_baseq, 1)) + row("Q30 Percentage", fixed(c.q30_percentage, 2) + "%");
        h += sec0 + "State Distribution" + sec1;
        h += row("Reference N", c.counts[0]) + row("Callable", c.counts[1]) + row("No Coverage", c.counts[2]) + row("Low Coverage", c.counts[3]) +
             row("Excessive Coverage", c.counts[4]) + row("Poor Mapping Quality", c.counts[5]);
        h += "</tbody></table>";
        if (c.has_plot) {
            const std::string plot = c.name + "_coverage.svg";
            h += "<figure class='coverage-plot'>\n                <img src=\"" + plot + "\" alt=\"Coverage distribution for " + c.name + "\" loading=\"lazy\">\n"
                 "                <figcaption>Coverage distribution for " + c.name + "</figcaption>\n            </figure>";
        }
        h += "</div>";
    }
    h += "</div></div>";
    h += footer;
    return h;
}

}  // namespace report
