// coverage_cli.cpp -- `decodingus-tools-b200 coverage`: the reference's `coverage` command on the B200 path.
//
// Host side of the drop-in, written in C++ because this image has no
// Rust toolchain; it only talks to the device through the C ABI of include/callable_loci_b200.h, exactly as a Rust
// host would.  Mirrors, with citations into /root/reference:
//   flags and defaults                     src/cli.rs:14-61
//   run_analysis / contig selection        src/api/coverage.rs:53-115,149-236 (ascending tid, largest non-chrM length)
//   process_single_contig                  src/callable_loci/mod.rs:44-147 (per-base loop replaced by the device)
//   build_coverage_export / natural order  src/callable_loci/report.rs:15-134,337-393
//   get_quality_stats                      src/callable_loci/profilers/contig_profiler.rs:123-158
//   summary.json (written in the CWD)      src/main.rs:67-69, src/export/formats/coverage.rs (field order)
//   detect_aligner / reference build       src/callable_loci/mod.rs:149-177, src/types.rs:100-147
//   BamStats sampler, platform inference, SVG plots, HTML report: report_writer.hpp (citations there)
// What rust-htslib did (BGZF inflate, BAM record decode, faidx) is done in bam_reader.hpp: multi-threaded zlib inflate of BGZF
// blocks, a record scan (sequential, or per selected contig through the .bai when there is one) and an in-memory FASTA contig load.
// The HTML page is wrapped in this tool's own header/footer unless --report-templates points at a reference checkout's
// src/callable_loci/templates (then the page is what the reference writes, byte for byte).
#include "../../../include/callable_loci_b200.h"
#include "bam_reader.hpp"
#include "report_writer.hpp"

#include <algorithm>
#include <atomic>
#include <charconv>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace {

[[noreturn]] void die(const std::string &m) { throw std::runtime_error(m); }
using namespace bamio;

inline uint64_t ticks() {
#if defined(__x86_64__)
    return __builtin_ia32_rdtsc();
#else
    return (uint64_t)std::chrono::steady_clock::now().time_since_epoch().count();
#endif
}

// ------------------------------------------------------------------------------------------------ report
struct ContigStats {            // ContigProfiler + CallableProfiler::contig_counts
    std::string name; uint64_t length = 0;
    uint64_t counts[6] = {0, 0, 0, 0, 0, 0};
    uint64_t n_covered = 0, sum_cov = 0, sum_bq = 0, sum_mapq = 0, qbases = 0, n_reads = 0;
};

size_t split_pos(const std::string &s) {
    for (size_t i = 0; i < s.size(); i++) if (isdigit((unsigned char)s[i]) || s[i] == 'X' || s[i] == 'Y' || s[i] == 'M') return i;
    return s.size();
}
void order_key(const std::string &suf, unsigned &cat, uint32_t &num) {
    const char *p = suf.c_str(); if (*p == '+') p++;
    bool ok = *p != 0; uint64_t v = 0;
    for (const char *q = p; *q; q++) { if (*q < '0' || *q > '9') { ok = false; break; } v = v * 10 + (uint64_t)(*q - '0'); if (v > 0xffffffffull) { ok = false; break; } }
    num = 0;
    if (ok) { cat = 0; num = (uint32_t)v; return; }
    cat = suf == "X" ? 1 : suf == "Y" ? 2 : (suf == "M" || suf == "MT") ? 3 : 4;
}
bool contig_less(const std::string &a, const std::string &b) {     // report.rs:339-383
    const size_t sa = split_pos(a), sb = split_pos(b);
    const std::string pa = a.substr(0, sa), pb = b.substr(0, sb);
    if (pa != pb) return pa < pb;
    unsigned ca, cb; uint32_t na, nb;
    order_key(a.substr(sa), ca, na); order_key(b.substr(sb), cb, nb);
    if (ca != cb) return ca < cb;
    if (ca == 0) return na < nb;
    return a.substr(sa) < b.substr(sb);
}

using report::fmt_f64;
using report::jstr;

std::string detect_aligner(const std::string &header_text) {       // callable_loci/mod.rs:149-177
    std::string h = header_text; for (auto &c : h) c = (char)tolower((unsigned char)c);
    auto has = [&](const char *s) { return h.find(s) != std::string::npos; };
    if (has("@pg\tid:bwa-mem2")) return "BWA-MEM2";
    if (has("@pg\tid:bwa")) return "BWA";
    if (has("@pg\tid:minimap2")) return "minimap2";
    if (has("@pg\tid:pbmm2")) return "pbmm2";
    if (has("@pg\tid:bowtie2")) return "Bowtie2";
    if (has("@pg\tid:star")) return "STAR";
    if (has("bwa")) return "BWA";
    if (has("minimap2")) return "minimap2";
    if (has("bowtie2")) return "Bowtie2";
    if (has("star")) return "STAR";
    return "Unknown";
}
std::string reference_build(const std::string &t) {                // types.rs:100-147
    auto has = [&](const char *s) { return t.find(s) != std::string::npos; };
    if (has("AS:GRCh38") || has("GCA_000001405.15")) return "GRCh38";
    if (has("AS:GRCh37") || has("GCA_000001405.1")) return "GRCh37";
    if (has("AS:CHM13") || has("GCA_009914755.4") || has("chm13") || has("CHM13") || has("t2t") || has("T2T")) return "T2T-CHM13v2.0";
    if (has("SN:chr1") && has("LN:248387328") && has("M5:e469247288ceb332aee524caec92bb22")) return "T2T-CHM13v2.0";
    if (has("SN:chr1") && has("LN:248956422")) return "GRCh38";
    if (has("SN:1") && has("LN:249250621")) return "GRCh37";
    return "Unknown";
}

// One page-locked column batch of admitted records (clb_host_alloc): what clb_push_reads copies from asynchronously.
struct PinBatch {
    size_t n = 0, n_cigar = 0, n_qual = 0, cap_n = 0, cap_c = 0, cap_q = 0;
    int32_t *pos = nullptr; uint16_t *flag = nullptr; uint8_t *mapq = nullptr;
    uint32_t *cigar_off = nullptr, *cigar = nullptr; uint64_t *qual_off = nullptr; uint8_t *qual = nullptr;
    template <class T> static void grow(T *&p, size_t used, size_t ncap) {
        T *q = (T *)clb_host_alloc(ncap * sizeof(T));
        if (!q) die("cannot allocate page-locked host memory");
        if (p) { memcpy(q, p, used * sizeof(T)); clb_host_free(p); }
        p = q;
    }
    void reserve(size_t rn, size_t rc, size_t rq) {
        if (rn > cap_n) { grow(pos, n, rn); grow(flag, n, rn); grow(mapq, n, rn); grow(cigar_off, n + 1, rn + 1); grow(qual_off, n + 1, rn + 1); cap_n = rn; }
        if (rc > cap_c) { grow(cigar, n_cigar, rc); cap_c = rc; }
        if (rq > cap_q) { grow(qual, n_qual, rq); cap_q = rq; }
        cigar_off[0] = 0; qual_off[0] = 0;
    }
    void clear() { n = n_cigar = n_qual = 0; }
    bool fits(const BamRecordView &r) const { return n + 1 <= cap_n && n_cigar + r.n_cigar <= cap_c && n_qual + (size_t)std::max(0, r.l_seq) <= cap_q; }
    void push(const BamRecordView &r) {
        const size_t lq = (size_t)std::max(0, r.l_seq);
        pos[n] = r.pos; flag[n] = r.flag; mapq[n] = r.mapq;
        memcpy(cigar + n_cigar, r.cigar, 4 * (size_t)r.n_cigar); n_cigar += r.n_cigar;
        memcpy(qual + n_qual, r.qual, lq); n_qual += lq;
        n++;
        cigar_off[n] = (uint32_t)n_cigar; qual_off[n] = n_qual;
    }
    void release() {
        for (void *p : {(void *)pos, (void *)flag, (void *)mapq, (void *)cigar_off, (void *)cigar, (void *)qual_off, (void *)qual}) clb_host_free(p);
        *this = PinBatch();
    }
};

// QNAMEs of the admitted records of one contig that reach >= 1 column, bucketed by hash: the exact distinct count
// (contig_profiler.rs:59-62) is a sort + compare per bucket, buckets in parallel.
struct NameStore {
    static constexpr unsigned kBuckets = 64;
    struct Ent { uint64_t hash; uint64_t off; uint32_t len; };
    std::vector<char> arena;
    std::vector<Ent> bucket[kBuckets];
    static uint64_t fnv(const char *p, size_t n) { uint64_t h = 1469598103934665603ull; for (size_t i = 0; i < n; i++) { h ^= (uint8_t)p[i]; h *= 1099511628211ull; } return h ^ (h >> 29); }
    void add(const char *p, uint32_t n) {
        const uint64_t h = fnv(p, n);
        bucket[h >> 58].push_back({h, arena.size(), n});
        arena.insert(arena.end(), p, p + n);
    }
    uint64_t count_unique(unsigned threads) {
        std::vector<uint64_t> cnt(kBuckets, 0);
        std::atomic<unsigned> next{0};
        auto work = [&] {
            for (;;) {
                const unsigned b = next.fetch_add(1);
                if (b >= kBuckets) break;
                auto &v = bucket[b];
                auto less = [&](const Ent &x, const Ent &y) {
                    if (x.hash != y.hash) return x.hash < y.hash;
                    const int c = memcmp(arena.data() + x.off, arena.data() + y.off, std::min(x.len, y.len));
                    return c != 0 ? c < 0 : x.len < y.len;
                };
                std::sort(v.begin(), v.end(), less);
                uint64_t n = 0;
                for (size_t i = 0; i < v.size(); i++)
                    if (i == 0 || v[i].hash != v[i - 1].hash || v[i].len != v[i - 1].len || memcmp(arena.data() + v[i].off, arena.data() + v[i - 1].off, v[i].len) != 0) n++;
                cnt[b] = n;
            }
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < std::max(1u, std::min(threads, kBuckets)); t++) th.emplace_back(work);
        work();
        for (auto &t : th) t.join();
        uint64_t n = 0; for (uint64_t c : cnt) n += c;
        return n;
    }
};

// decoder thread -> device thread
struct Msg { PinBatch *batch = nullptr; bool end_of_contig = false; NameStore *names = nullptr; uint64_t admitted = 0; std::string error; };
template <class T> class Channel {
  public:
    void put(T v) { { std::lock_guard<std::mutex> g(m_); q_.push_back(std::move(v)); } cv_.notify_one(); }
    T take(double *waited_s = nullptr) {
        std::unique_lock<std::mutex> g(m_);
        const auto t0 = std::chrono::steady_clock::now();
        cv_.wait(g, [this] { return !q_.empty(); });
        if (waited_s) *waited_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        T v = std::move(q_.front()); q_.pop_front();
        return v;
    }
  private:
    std::mutex m_; std::condition_variable cv_; std::deque<T> q_;
};

struct Options {
    std::string bam, reference, out_bed = "callable_regions.bed", summary = "summary.html", templates;
    std::vector<std::string> contigs; bool have_contigs = false;
    clb_options o{4, 500, 10, 10, 20, 1, 0, 0.1};
    unsigned threads = std::max(1u, std::thread::hardware_concurrency());
    int device = 0;
    bool verbose = false, progress = false; std::string timing_json;
};

void usage() {
    fprintf(stderr, "Usage: decodingus-tools-b200 coverage <BAM_FILE> -r <REFERENCE> [-o callable_regions.bed] [-s summary.html] [-L contig]...\n"
                    "       [--min-depth 4] [--max-depth 500] [--min-mapping-quality 10] [--min-base-quality 20]\n"
                    "       [--min-depth-for-low-mapq 10] [--max-low-mapq 1] [--max-low-mapq-fraction 0.1] [--threads N] [--device D]\n"
                    "       [--report-templates DIR] [--verbose] [--progress] [--timing-json FILE]\n");
    exit(2);
}

int run(int argc, char **argv) {
    if (argc < 3 || strcmp(argv[1], "coverage") != 0) usage();
    Options opt;
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() -> std::string { if (i + 1 >= argc) usage(); return argv[++i]; };
        if (a == "-r" || a == "--reference") opt.reference = val();
        else if (a == "-o" || a == "--output") opt.out_bed = val();
        else if (a == "-s" || a == "--summary") opt.summary = val();
        else if (a == "-L" || a == "--contig") { opt.contigs.push_back(val()); opt.have_contigs = true; }
        else if (a == "--min-depth") opt.o.min_depth = (uint32_t)std::stoul(val());
        else if (a == "--max-depth") opt.o.max_depth = (uint32_t)std::stoul(val());
        else if (a == "--min-mapping-quality") opt.o.min_mapping_quality = (uint8_t)std::stoul(val());
        else if (a == "--min-base-quality") opt.o.min_base_quality = (uint8_t)std::stoul(val());
        else if (a == "--min-depth-for-low-mapq") opt.o.min_depth_for_low_mapq = (uint32_t)std::stoul(val());
        else if (a == "--max-low-mapq") opt.o.max_low_mapq = (uint8_t)std::stoul(val());
        else if (a == "--max-low-mapq-fraction") opt.o.max_low_mapq_fraction = std::stod(val());
        else if (a == "--threads") opt.threads = (unsigned)std::stoul(val());
        else if (a == "--device") opt.device = std::stoi(val());
        else if (a == "--report-templates") opt.templates = val();
        else if (a == "--verbose") opt.verbose = true;
        else if (a == "--progress") opt.progress = true;
        else if (a == "--timing-json") opt.timing_json = val();
        else if (!a.empty() && a[0] != '-' && opt.bam.empty()) opt.bam = a;
        else usage();
    }
    if (opt.bam.empty() || opt.reference.empty()) usage();

    BamReader bam(opt.bam, opt.threads);
    const BamHeader &H = bam.header();
    const auto fai = load_fai(opt.reference);

    // initialize_contig_stats + validate_contig_selection (api/coverage.rs:149-204)
    std::set<std::string> selected(opt.contigs.begin(), opt.contigs.end());
    std::map<int32_t, ContigStats> stats;
    for (size_t tid = 0; tid < H.names.size(); tid++) {
        if (opt.have_contigs && !selected.count(H.names[tid])) continue;
        ContigStats s; s.name = H.names[tid]; s.length = H.lens[tid];
        stats[(int32_t)tid] = s;
    }
    if (opt.have_contigs && stats.empty()) {
        std::string l; for (size_t i = 0; i < opt.contigs.size(); i++) l += (i ? ", " : "") + opt.contigs[i];
        die("None of the specified contigs (" + l + ") were found in the BAM file");
    }
    uint32_t largest = 0;                                               // initialize_counter (api/coverage.rs:206-219)
    for (auto &kv : stats) if (kv.second.name != "chrM") largest = std::max<uint32_t>(largest, (uint32_t)kv.second.length);

    char err[512];
    clb_ctx *ctx = clb_create(opt.device, &opt.o, err, sizeof err);
    if (!ctx) die(std::string("GPU context: ") + err);
    clb_bed_writer *bed = clb_bed_writer_open(opt.out_bed.c_str(), largest);
    if (!bed) die("Failed to create CallableProfiler: cannot create " + opt.out_bed);
    auto check = [&](int rc, const char *what) { if (rc != 0) die(std::string("Error processing contig: ") + what + ": " + clb_last_error(ctx)); };

    // BamStats: its own pass over the first 10 000 records of the file, primary alignments only (bam_stats.rs:44-76)
    report::BamStats bs;
    {
        BamReader head(opt.bam, std::min(opt.threads, 4u));
        BamRecordView r;
        while (!bs.full() && head.next(r)) bs.add_record(std::string(r.qname, r.l_qname ? r.l_qname - 1 : 0), r.flag, (uint64_t)std::max(0, r.l_seq));
    }
    // With a contig selection and a .bai next to the BAM the reader jumps to each contig (what bam.fetch does, mod.rs:54);
    // otherwise the coordinate-sorted file is scanned once from the start.
    BaiIndex bai;
    const bool indexed = opt.have_contigs && load_bai(opt.bam, bai) && bai.first.size() == H.names.size();
    // coverage plots land next to the BED file (callable_profiler.rs:24-27,80-83)
    const size_t slash = opt.out_bed.find_last_of('/');
    const std::string out_dir = slash == std::string::npos ? "" : opt.out_bed.substr(0, slash + 1);

    // ------------------------------------------------------------------------------------------ the pipeline
    // decoder thread: BGZF read-ahead + inflate pool (bam_reader.hpp) -> record scan -> htslib admission, one record at a
    //                 time (clb_admitter_*) -> admitted records packed into page-locked column batches + QNAME store
    // this thread   : clb_begin_contig / clb_push_reads per batch (copies overlap the kernels of earlier windows and the
    //                 decoding of the next batch) / clb_finish_contig, BED text, plots
    const uint32_t maxcnt = opt.o.max_depth > 0 ? opt.o.max_depth : 500;
    const auto wall0 = std::chrono::steady_clock::now();
    auto secs = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count(); };
    // batches of about a million reads for real inputs, small ones for small files (page-locking memory is not free)
    size_t kBatchReads = 1u << 20;
    {
        FILE *fsz = fopen(opt.bam.c_str(), "rb");
        if (fsz) { fseeko(fsz, 0, SEEK_END); const off_t sz = ftello(fsz); fclose(fsz); if (sz < (off_t)(256u << 20)) kBatchReads = std::max<size_t>(4096, (size_t)sz / 64); }
    }
    constexpr int kBatches = 3;
    PinBatch pool[kBatches];
    Channel<PinBatch *> free_batches;
    Channel<Msg> ready;
    for (auto &b : pool) { b.reserve(kBatchReads, 2 * kBatchReads, 160 * kBatchReads); free_batches.put(&b); }
    double decode_wait_s = 0, decode_total_s = 0, admit_s = 0;
    uint64_t admit_ticks = 0;                      // per-record timing with the cycle counter (a clock call per record would cost more than the admission)
    std::vector<int32_t> tids; std::map<int32_t, uint32_t> lens;
    for (auto &kv : stats) { tids.push_back(kv.first); lens[kv.first] = (uint32_t)kv.second.length; }
    std::thread decoder([&] {
        const auto t_start = std::chrono::steady_clock::now();
        const uint64_t tick_start = ticks();
        try {
            BamRecordView rec; bool have = indexed ? false : bam.next(rec);
            for (const int32_t tid : tids) {                              // ascending tid (api/coverage.rs:229-235)
                const uint32_t clen = lens.at(tid);
                if (indexed) {
                    have = bai.first[(size_t)tid] != UINT64_MAX;
                    if (have) { bam.seek(bai.first[(size_t)tid]); have = bam.next(rec); }
                }
                while (have && (rec.tid < tid && rec.tid >= 0)) have = bam.next(rec);   // records of contigs that were not selected
                clb_admitter *adm = clb_admitter_new(tid, maxcnt);
                NameStore *names = new NameStore();
                PinBatch *cur = nullptr; uint64_t admitted = 0;
                while (have && rec.tid == tid) {
                    const uint64_t ta = ticks();
                    const int k = clb_admitter_push(adm, rec.pos, rec.flag, rec.cigar, rec.n_cigar);
                    admit_ticks += ticks() - ta;
                    if (k < 0) { clb_admitter_free(adm); die("Error processing contig: records are not coordinate sorted"); }
                    if (k == 1) {
                        if (!cur) { cur = free_batches.take(&decode_wait_s); cur->clear(); }
                        if (!cur->fits(rec)) {
                            if (cur->n) { Msg m; m.batch = cur; ready.put(std::move(m)); cur = free_batches.take(&decode_wait_s); cur->clear(); }
                            if (!cur->fits(rec)) cur->reserve(std::max<size_t>(cur->cap_n, 1), std::max<size_t>(cur->cap_c, 2 * (size_t)rec.n_cigar),
                                                              std::max<size_t>(cur->cap_q, 2 * (size_t)std::max(0, rec.l_seq)));
                        }
                        cur->push(rec); admitted++;
                        uint64_t span = 0;
                        for (uint32_t c = 0; c < rec.n_cigar; c++) { const uint32_t op = rec.cigar[c] & 15; if ((0x18du >> op) & 1u) span += rec.cigar[c] >> 4; }
                        if (span && (uint32_t)rec.pos < clen) names->add(rec.qname, rec.l_qname ? rec.l_qname - 1 : 0);
                    }
                    have = bam.next(rec);
                }
                clb_admitter_free(adm);
                Msg m; m.batch = cur; m.end_of_contig = true; m.names = names; m.admitted = admitted;
                ready.put(std::move(m));
            }
        } catch (const std::exception &e) {
            Msg m; m.error = e.what(); m.end_of_contig = true; ready.put(std::move(m));
        }
        decode_total_s = secs(t_start);
        admit_s = decode_total_s > 0 ? (double)admit_ticks * decode_total_s / (double)std::max<uint64_t>(1, ticks() - tick_start) : 0.0;
    });

    double device_ms = 0, h2d_ms = 0, bed_s = 0, names_s = 0, ref_s = 0, device_wait_s = 0;
    uint64_t total_admitted = 0, total_cells = 0;
    std::string failure;
    // --progress: the reference API's ProgressEvent stream (api/mod.rs:12-18, emitted by CoverageAnalyzer::analyze at
    // api/coverage.rs:40-48), one serde-style JSON object per line on stderr; Progress counts finished contigs
    auto progress = [&](const char *kind, const std::string &extra) {
        if (opt.progress) fprintf(stderr, "{\"%s\":{\"task\":\"Coverage Analysis\"%s}}\n", kind, extra.c_str());
    };
    progress("Started", "");
    size_t contigs_done = 0;
    for (const int32_t tid : tids) {
        ContigStats &st = stats[tid];
        const uint32_t clen = (uint32_t)st.length;
        const auto tr = std::chrono::steady_clock::now();
        std::vector<uint8_t> ref;
        auto fe = fai.find(st.name);
        if (fe != fai.end()) ref = load_contig(opt.reference, fe->second);   // missing contig / short sequence reads as 'N'
        ref_s += secs(tr);
        check(clb_begin_contig(ctx, tid, st.name.c_str(), clen, ref.data(), ref.size(), 0, largest, 0, clen, 0), "begin");
        NameStore *names = nullptr; uint64_t admitted = 0;
        for (;;) {
            Msg m = ready.take(&device_wait_s);
            if (!m.error.empty()) { failure = m.error; break; }
            if (m.batch) {
                if (m.batch->n) {
                    clb_read_batch rb{m.batch->n, m.batch->n_cigar, m.batch->n_qual, m.batch->pos, m.batch->flag, m.batch->mapq,
                                      m.batch->cigar_off, m.batch->cigar, m.batch->qual_off, m.batch->qual};
                    check(clb_push_reads(ctx, &rb), "push");
                    check(clb_wait_uploads(ctx), "upload");                  // the batch's buffers go back to the decoder
                }
                free_batches.put(m.batch);
            }
            if (m.end_of_contig) { names = m.names; admitted = m.admitted; break; }
        }
        if (!failure.empty()) break;
        clb_contig_result res{};
        check(clb_finish_contig(ctx, &res), "finish");
        device_ms += res.kernel_ms; h2d_ms += res.h2d_ms;
        const auto tn = std::chrono::steady_clock::now();
        st.n_reads = names ? names->count_unique(opt.threads) : 0;
        delete names;
        names_s += secs(tn);
        for (int s = 0; s < 6; s++) st.counts[s] = res.state_counts[s];
        st.n_covered = res.n_covered_bases; st.sum_cov = res.summed_coverage; st.sum_bq = res.summed_baseq;
        st.sum_mapq = res.summed_mapq; st.qbases = res.quality_bases;
        total_admitted += admitted; total_cells += res.summed_coverage;
        const auto tb = std::chrono::steady_clock::now();
        std::vector<uint32_t> bins(res.bins, res.bins + 3 * (size_t)res.n_bins);
        int has_bins = 0;
        if (clb_bed_writer_add_contig(bed, st.name.c_str(), clen, res.intervals, res.n_intervals, bins.data(), res.n_bins, res.stride, &has_bins) != 0)
            die("Error processing contig: interval list does not tile the contig");
        if (has_bins && res.stride) {                                        // finish_contig, callable_profiler.rs:64-86
            const std::string svg = report::render_coverage_svg(st.name, clen, res.stride, bins.data(), res.n_bins);
            const std::string path = out_dir + st.name + "_coverage.svg";
            FILE *fs = fopen(path.c_str(), "wb");
            if (!fs) die("Error processing contig: cannot create " + path);
            fwrite(svg.data(), 1, svg.size(), fs); fclose(fs);
        }
        bed_s += secs(tb);
        if (opt.verbose)
            fprintf(stderr, "%s: %llu admitted reads, %llu cells, %.2f ms on device\n", st.name.c_str(), (unsigned long long)admitted,
                    (unsigned long long)res.summed_coverage, res.kernel_ms);
        progress("Progress", ",\"current\":" + std::to_string(++contigs_done) + ",\"total\":" + std::to_string(tids.size()));
    }
    if (!failure.empty()) {
        // let the decoder run dry so that it can be joined
        std::thread drain([&] { for (;;) { Msg m = ready.take(); if (m.batch) free_batches.put(m.batch); delete m.names; if (!m.error.empty()) break; } });
        drain.detach();
        decoder.detach();
        progress("Error", ",\"error\":\"" + failure + "\"");
        die(failure);
    }
    decoder.join();
    for (auto &b : pool) b.release();
    const double wall_s = secs(wall0);
    fprintf(stderr, "coverage: %zu contigs, %llu admitted reads, %llu aligned bases | wall %.3f s | host decode %.3f s (BGZF inflate %.3f s on %u threads, "
                    "admission %.3f s inline, waiting for a free batch %.3f s) | device %.3f s kernels, %.3f s H2D | reference load %.3f s | "
                    "unique names %.3f s | BED + plots %.3f s | device thread waited %.3f s for the decoder\n",
            tids.size(), (unsigned long long)total_admitted, (unsigned long long)total_cells, wall_s, decode_total_s - decode_wait_s,
            bam.inflate_seconds(), opt.threads, admit_s, decode_wait_s, device_ms * 1e-3, h2d_ms * 1e-3, ref_s, names_s, bed_s, device_wait_s);
    if (!opt.timing_json.empty()) {
        FILE *ft = fopen(opt.timing_json.c_str(), "wb");
        if (ft) {
            fprintf(ft, "{\"contigs\": %zu, \"admitted_reads\": %llu, \"aligned_bases\": %llu, \"wall_s\": %.6f, \"decode_s\": %.6f, \"inflate_s\": %.6f, "
                        "\"threads\": %u, \"admission_s\": %.6f, \"decoder_waited_for_batch_s\": %.6f, \"device_kernels_s\": %.6f, \"h2d_s\": %.6f, "
                        "\"reference_load_s\": %.6f, \"unique_names_s\": %.6f, \"bed_and_plots_s\": %.6f, \"device_thread_waited_s\": %.6f}\n",
                    tids.size(), (unsigned long long)total_admitted, (unsigned long long)total_cells, wall_s, decode_total_s - decode_wait_s, bam.inflate_seconds(),
                    opt.threads, admit_s, decode_wait_s, device_ms * 1e-3, h2d_ms * 1e-3, ref_s, names_s, bed_s, device_wait_s);
            fclose(ft);
        }
    }
    if (clb_bed_writer_close(bed) != 0) die("failed to write " + opt.out_bed);
    clb_destroy(ctx);

    // build_coverage_export (report.rs:15-134)
    std::vector<const ContigStats *> order;
    for (auto &kv : stats) order.push_back(&kv.second);
    std::stable_sort(order.begin(), order.end(), [](const ContigStats *a, const ContigStats *b) { return contig_less(a->name, b->name); });
    uint64_t total_bases = 0, callable = 0, q30_bases = 0, total_qpos = 0, total_unique = 0;
    double total_depth = 0, total_mapq = 0, total_baseq = 0;
    std::string cj;
    std::vector<report::ContigRow> rows;
    auto exists = [](const std::string &p) { FILE *f = fopen(p.c_str(), "rb"); if (f) fclose(f); return f != nullptr; };
    for (size_t i = 0; i < order.size(); i++) {
        const ContigStats &s = *order[i];
        const double amq = s.qbases ? (double)s.sum_mapq / (double)s.qbases : 0.0, abq = s.qbases ? (double)s.sum_bq / (double)s.qbases : 0.0;
        const double q30 = s.qbases ? (abq >= 30.0 ? 100.0 : abq < 20.0 ? 0.0 : ((abq - 20.0) / 10.0) * 100.0) : 0.0;
        const double covp = s.length ? ((double)s.n_covered / (double)s.length) * 100.0 : 0.0;
        const double adep = s.n_covered ? (double)s.sum_cov / (double)s.n_covered : 0.0;
        total_bases += s.length; callable += s.counts[CLB_CALLABLE];
        total_depth += adep * (double)s.length; total_mapq += amq * (double)s.length; total_baseq += abq * (double)s.length;
        q30_bases += (uint64_t)(q30 / 100.0 * (double)s.length); total_qpos += s.length; total_unique += (uint32_t)s.n_reads;
        report::ContigRow row{s.name, s.length, s.n_reads, s.n_covered, covp, adep, amq, abq, q30, {0, 0, 0, 0, 0, 0},
                              exists(s.name + "_coverage.svg")};   // path relative to the CWD, as report.rs:318-319 tests it
        for (int k = 0; k < 6; k++) row.counts[k] = s.counts[k];
        rows.push_back(row);
        cj += std::string(i ? ",\n" : "") + "      {\n        \"name\": " + jstr(s.name) + ",\n        \"length\": " + std::to_string(s.length) +
              ",\n        \"unique_reads\": " + std::to_string(s.n_reads) + ",\n        \"coverage_percent\": " + fmt_f64(covp) +
              ",\n        \"average_depth\": " + fmt_f64(adep) + ",\n        \"covered_bases\": " + std::to_string(s.n_covered) +
              ",\n        \"total_bases\": " + std::to_string(s.length) + ",\n        \"quality_stats\": {\n          \"average_mapq\": " + fmt_f64(amq) +
              ",\n          \"average_baseq\": " + fmt_f64(abq) + ",\n          \"q30_percentage\": " + fmt_f64(q30) +
              "\n        },\n        \"state_distribution\": {\n          \"ref_n\": " + std::to_string(s.counts[0]) + ",\n          \"callable\": " +
              std::to_string(s.counts[1]) + ",\n          \"no_coverage\": " + std::to_string(s.counts[2]) + ",\n          \"low_coverage\": " +
              std::to_string(s.counts[3]) + ",\n          \"excessive_coverage\": " + std::to_string(s.counts[4]) +
              ",\n          \"poor_mapping_quality\": " + std::to_string(s.counts[5]) + "\n        }\n      }";
    }
    const double avg_depth = total_bases ? total_depth / (double)total_bases : 0.0;
    const double call_pct = total_bases ? ((double)callable / (double)total_bases) * 100.0 : 0.0;
    // collect_coverage_plots (api/coverage.rs:262-275): plots found in the CWD; the reference lists them in HashMap order, here by tid
    std::string plots;
    for (auto &kv : stats) {
        const std::string p = kv.second.name + "_coverage.svg";
        if (exists(p)) plots += std::string(plots.empty() ? "\n      " : ",\n      ") + jstr(p);
    }
    if (!plots.empty()) plots += "\n    ";
    std::string js = "{\n  \"export\": {\n    \"summary\": {\n      \"aligner\": " + jstr(detect_aligner(H.text)) + ",\n      \"reference_build\": " +
        jstr(reference_build(H.text)) + ",\n      \"sequencing_platform\": " + jstr(bs.infer_platform()) + ",\n      \"read_length\": " + std::to_string(bs.average_read_length()) +
        ",\n      \"total_bases\": " + std::to_string(total_bases) + ",\n      \"callable_bases\": " + std::to_string(callable) +
        ",\n      \"callable_percentage\": " + fmt_f64(call_pct) + ",\n      \"average_depth\": " + fmt_f64(avg_depth) + ",\n      \"contigs_analyzed\": " +
        std::to_string(stats.size()) + "\n    },\n    \"contigs\": [" + (order.empty() ? "" : "\n" + cj + "\n    ") + "],\n    \"quality_metrics\": {\n      \"average_mapq\": " +
        fmt_f64(total_qpos ? total_mapq / (double)total_qpos : 0.0) + ",\n      \"average_baseq\": " + fmt_f64(total_qpos ? total_baseq / (double)total_qpos : 0.0) +
        ",\n      \"q30_percentage\": " + fmt_f64(total_qpos ? ((double)q30_bases / (double)total_qpos) * 100.0 : 0.0) + "\n    },\n    \"total_unique_reads\": " +
        std::to_string(total_unique) + "\n  },\n  \"files\": {\n    \"bed_file\": " + jstr(opt.out_bed) + ",\n    \"summary_html\": " + jstr(opt.summary) +
        ",\n    \"coverage_plots\": [" + plots + "]\n  }\n}";
    FILE *fj = fopen("summary.json", "wb");                                  // main.rs:68: always in the CWD
    if (!fj) die("cannot create summary.json");
    fwrite(js.data(), 1, js.size(), fj); fclose(fj);

    // write_html_report (report.rs:136-160)
    auto slurp = [](const std::string &p) { std::ifstream f(p, std::ios::binary); if (!f) die("cannot read " + p); return std::string((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>()); };
    const std::string header = opt.templates.empty() ? report::default_header() : slurp(opt.templates + "/report_header.html");
    const std::string footer = opt.templates.empty() ? report::default_footer() : slurp(opt.templates + "/report_footer.html");
    report::Summary sm{reference_build(H.text), detect_aligner(H.text), bs.infer_platform(), bs.average_read_length(), total_unique, total_bases, callable,
                       (uint64_t)stats.size(), bs.max_samples, call_pct, avg_depth, total_qpos ? total_mapq / (double)total_qpos : 0.0,
                       total_qpos ? total_baseq / (double)total_qpos : 0.0};
    const std::string html = report::render_html_report(sm, rows, header, footer);
    FILE *fh = fopen(opt.summary.c_str(), "wb");
    if (!fh) die("cannot create " + opt.summary);
    fwrite(html.data(), 1, html.size(), fh); fclose(fh);
    progress("Completed", "");
    return 0;
}

}  // namespace

int main(int argc, char **argv) {
    try {
        return run(argc, argv);
    } catch (const std::exception &e) {      // every failure: message + exit code 1, the partially written BED stays behind
        fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
}
