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

// ------------------------------------------------------------------------------------------------ BAM / FASTA reader
#include "../../decodingustools_b200/csrc/host/bam_reader.hpp"

extern "C" {

// Scans the whole file into caller-provided columns.  Returns 0, or 1 with a message in err.
int shim_bam_scan(const char *path, unsigned threads, uint64_t cap_reads, uint64_t cap_cigar, uint64_t cap_qual, int32_t *tid, int32_t *pos,
                  uint16_t *flag, uint8_t *mapq, uint32_t *cigar_off, uint32_t *cigar, uint64_t *qual_off, uint8_t *qual, char *names,
                  size_t names_cap, char *hdr_text, size_t hdr_cap, char *ref_names, size_t ref_names_cap, uint32_t *ref_lens, uint32_t *n_ref,
                  uint64_t *n_reads, char *err, size_t err_cap) {
    try {
        bamio::BamReader rd(path, threads);
        const bamio::BamHeader &h = rd.header();
        put(h.text, hdr_text, hdr_cap);
        std::string rn; for (auto &n : h.names) { rn += n; rn.push_back('\n'); }
        put(rn, ref_names, ref_names_cap);
        *n_ref = (uint32_t)h.names.size();
        for (size_t i = 0; i < h.lens.size(); i++) ref_lens[i] = h.lens[i];
        bamio::BamRecordView r; uint64_t n = 0, nc = 0, nq = 0; size_t no = 0;
        cigar_off[0] = 0; qual_off[0] = 0;
        while (rd.next(r)) {
            const uint64_t lq = (uint64_t)std::max(0, r.l_seq), ln = r.l_qname ? r.l_qname - 1 : 0;
            if (n >= cap_reads || nc + r.n_cigar > cap_cigar || nq + lq > cap_qual || no + ln + 1 > names_cap) throw std::runtime_error("shim buffers too small");
            tid[n] = r.tid; pos[n] = r.pos; flag[n] = r.flag; mapq[n] = r.mapq;
            memcpy(cigar + nc, r.cigar, 4 * (size_t)r.n_cigar); nc += r.n_cigar; cigar_off[n + 1] = (uint32_t)nc;
            memcpy(qual + nq, r.qual, lq); nq += lq; qual_off[n + 1] = nq;
            memcpy(names + no, r.qname, ln); names[no + ln] = '\n'; no += ln + 1;
            n++;
        }
        names[no < names_cap ? no : names_cap - 1] = 0;
        *n_reads = n;
        return 0;
    } catch (const std::exception &e) { put(e.what(), err, err_cap); return 1; }
}

int64_t shim_load_contig(const char *fasta, const char *name, uint8_t *out, uint64_t cap, char *err, size_t err_cap) {
    try {
        const auto fai = bamio::load_fai(fasta);
        const auto it = fai.find(name);
        if (it == fai.end()) return -1;
        const auto seq = bamio::load_contig(fasta, it->second);
        if (seq.size() > cap) throw std::runtime_error("shim buffer too small");
        memcpy(out, seq.data(), seq.size());
        return (int64_t)seq.size();
    } catch (const std::exception &e) { put(e.what(), err, err_cap); return -2; }
}

}  // extern "C"

extern "C" size_t shim_fmt_f64(double v, char *out, size_t cap) { return put(report::fmt_f64(v), out, cap); }
extern "C" size_t shim_jstr(const char *s, char *out, size_t cap) { return put(report::jstr(s), out, cap); }

// Records of one contig through the BAI index (seek + scan until the tid changes): returns the record count and the
// sum of their positions, -1 without an index, -2 on error.
extern "C" int64_t shim_bam_fetch_tid(const char *path, int32_t tid, int64_t *pos_sum, uint64_t *first_voffset, char *err, size_t err_cap) {
    try {
        bamio::BaiIndex bai;
        if (!bamio::load_bai(path, bai)) return -1;
        bamio::BamReader rd(path, 2);
        *first_voffset = bai.first.at((size_t)tid);
        if (*first_voffset == UINT64_MAX) return 0;
        rd.seek(*first_voffset);
        bamio::BamRecordView r; int64_t n = 0; *pos_sum = 0;
        while (rd.next(r) && r.tid == tid) { n++; *pos_sum += r.pos; }
        return n;
    } catch (const std::exception &e) { put(e.what(), err, err_cap); return -2; }
}
