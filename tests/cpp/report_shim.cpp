// Test-only shim: exposes csrc/host/report_writer.hpp (the C++ the CLI uses) through a tiny C ABI so that the CPU test
// suite can compare it with the Python mirror (decodingustools_b200/bam_stats.py, report.py) without a GPU.
#include "../../decodingustools_b200/csrc/host/report_writer.hpp"

#include <cstring>

static size_t put(const std::string &s, char *out, size_t cap) {
    if (out && cap) { const size_t n = s.size() < cap - 1 ? s.size() : cap - 1; memcpy(out, s.data(), n); out[n] = 0; }
    return s.size();
}

extern "C" {

int shim_detect_platform(const char *qname) { return (int)report::detect_platform_from_qname(qname); }

// records: n NUL-terminated names back to back; flags[n]; seq_len[n].  Returns the inferred platform string.
size_t shim_bam_stats(const char *names, const uint16_t *flags, const uint64_t *seq_len, uint64_t n, uint64_t max_samples,
                      uint64_t *read_count, uint64_t *avg_len, int *primary, char *out, size_t cap) {
    report::BamStats bs; bs.max_samples = max_samples;
    const char *p = names;
    for (uint64_t i = 0; i < n; i++) { const std::string q(p); p += q.size() + 1; bs.add_record(q, flags[i], seq_len[i]); }
    *read_count = bs.read_count; *avg_len = bs.average_read_length(); *primary = (int)bs.primary_platform();
    return put(bs.infer_platform(), out, cap);
}

size_t shim_svg(const char *contig, uint32_t contig_length, uint32_t stride, const uint32_t *bins, uint32_t n_bins, char *out, size_t cap) {
    return put(report::render_coverage_svg(contig, contig_length, stride, bins, n_bins), out, cap);
}

// one summary + n contig rows -> the page
size_t shim_html(const char *reference_build, const char *aligner, const char *platform, const uint64_t *su /* read_length, total_unique_reads,
                 total_bases, callable_bases, contigs_analyzed, max_samples */, const double *sd /* callable_percentage, average_depth,
                 average_mapq, average_baseq */, uint64_t n, const char *names, const uint64_t *cu /* per contig: length, unique_reads,
                 covered_bases, counts[6], has_plot */, const double *cd /* per contig: coverage_percent, average_depth, average_mapq,
                 average_baseq, q30_percentage */, const char *header, const char *footer, char *out, size_t cap) {
    report::Summary s{reference_build, aligner, platform, su[0], su[1], su[2], su[3], su[4], su[5], sd[0], sd[1], sd[2], sd[3]};
    std::vector<report::ContigRow> rows;
    const char *p = names;
    for (uint64_t i = 0; i < n; i++) {
        const std::string nm(p); p += nm.size() + 1;
        const uint64_t *u = cu + 10 * i; const double *d = cd + 5 * i;
        report::ContigRow r{nm, u[0], u[1], u[2], d[0], d[1], d[2], d[3], d[4], {u[3], u[4], u[5], u[6], u[7], u[8]}, u[9] != 0};
        rows.push_back(r);
    }
    return put(report::render_html_report(s, rows, header ? header : report::default_header(), footer ? footer : report::default_footer()), out, cap);
}

const char *shim_default_header() { return report::default_header(); }
const char *shim_default_footer() { return report::default_footer(); }

}  // extern "C"
