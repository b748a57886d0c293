// clb_host.cpp -- host half of include/callable_loci_b200.h (no GPU needed).
//
//   clb_admit_reads        htslib bam_plp_push admission as configured by /root/reference/src/callable_loci/mod.rs:55-60
//   clb_bed_writer_*       CallableProfiler::write_state / finish_contig incl. the cross-contig quirks,
//                          /root/reference/src/callable_loci/profilers/callable_profiler.rs:39-87,122-155
//   clb_stitch_intervals   region-shard stitching (multi-GPU)
#include "../../include/callable_loci_b200.h"

#include <cstdio>
#include <cstring>
#include <functional>
#include <queue>
#include <string>
#include <vector>

static const char *kStateName[6] = {"REF_N", "CALLABLE", "NO_COVERAGE", "LOW_COVERAGE", "EXCESSIVE_COVERAGE", "POOR_MAPPING_QUALITY"};

extern "C" int clb_admit_reads(int32_t tid, uint32_t maxcnt, uint64_t n_reads, const int32_t *pos, const uint16_t *flag,
                               const uint32_t *cigar_off, const uint32_t *cigar, uint8_t *keep) {
    // Closed form of the iterator mechanics: a record is dropped iff it is not the
    // first record seen at its start position and the number of live nodes (admitted records whose
    // end >= pos) has reached maxcnt.  The iterator starts at (tid 0, pos 0), so the very first record of
    // tid 0 at pos 0 counts as "not first".  A zero-span record that is not first is never retained.
    std::priority_queue<long long, std::vector<long long>, std::greater<long long>> live;   // min-heap of ends
    long long last_pos = (tid == 0) ? 0 : -1;
    bool any = false;
    long long prev = -1;
    for (uint64_t i = 0; i < n_reads; i++) {
        keep[i] = 0;
        if (flag[i] & 0x4) continue;
        const long long p = pos[i];
        if (p < prev) return CLB_E_INPUT;
        prev = p;
        long long span = 0;
        for (uint32_t c = cigar_off[i]; c < cigar_off[i + 1]; c++) {
            const uint32_t op = cigar[c] & 15u;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += cigar[c] >> 4;
        }
        const long long end = p + span;
        const bool same = any ? (p == last_pos) : (tid == 0 && p == 0);
        any = true;
        while (!live.empty() && live.top() < p) live.pop();
        if (same) {
            if (live.size() >= (size_t)maxcnt) continue;
            if (end > p) { live.push(end); keep[i] = 1; }
        } else {
            live.push(end); keep[i] = 1;
        }
        last_pos = p;
    }
    return CLB_OK;
}

extern "C" int clb_compact_reads(const clb_read_batch *in, const uint8_t *keep, int32_t *pos, uint16_t *flag, uint8_t *mapq,
                                 uint32_t *cigar_off, uint32_t *cigar, uint64_t *qual_off, uint8_t *qual, clb_read_batch *out) {
    if (!in || !keep || !out) return CLB_E_INVALID;
    uint64_t n = 0, nc = 0, nq = 0;
    cigar_off[0] = 0; qual_off[0] = 0;
    for (uint64_t i = 0; i < in->n_reads; i++) {
        if (!keep[i]) continue;
        const uint32_t c0 = in->cigar_off[i], c1 = in->cigar_off[i + 1];
        const uint64_t q0 = in->qual_off[i], q1 = in->qual_off[i + 1];
        pos[n] = in->pos[i]; flag[n] = in->flag[i]; mapq[n] = in->mapq[i];
        memcpy(cigar + nc, in->cigar + c0, (size_t)(c1 - c0) * 4); nc += c1 - c0;
        memcpy(qual + nq, in->qual + q0, (size_t)(q1 - q0)); nq += q1 - q0;
        n++;
        cigar_off[n] = (uint32_t)nc; qual_off[n] = nq;
    }
    out->n_reads = n; out->n_cigar = nc; out->n_qual = nq;
    out->pos = pos; out->flag = flag; out->mapq = mapq; out->cigar_off = cigar_off; out->cigar = cigar; out->qual_off = qual_off; out->qual = qual;
    return CLB_OK;
}

extern "C" uint64_t clb_stitch_intervals(const clb_interval *const *shards, const uint64_t *n_per_shard, uint32_t n_shards,
                                         clb_interval *out) {
    uint64_t n = 0;
    for (uint32_t s = 0; s < n_shards; s++) {
        for (uint64_t i = 0; i < n_per_shard[s]; i++) {
            clb_interval iv = shards[s][i];
            if (n > 0 && i == 0 && iv.soft_start && out[n - 1].state == iv.state && out[n - 1].end == iv.start) {
                out[n - 1].end = iv.end;                     // same run continues across the shard seam
                continue;
            }
            iv.soft_start = 0;
            out[n++] = iv;
        }
    }
    return n;
}

struct clb_bed_writer {
    FILE *fp = nullptr;
    std::string mem;
    bool in_memory = false;
    uint32_t largest = 0;
    bool have_pending = false;     // CallableProfiler::current_state survives finish_contig (quirk Q1)
    std::string pend_name;
    clb_interval pend{};
    std::vector<char> buf;
    void put(const char *s, size_t n) {
        if (in_memory) mem.append(s, n);
        else {
            if (buf.size() + n > (1u << 20)) flush();
            buf.insert(buf.end(), s, s + n);
        }
    }
    void flush() { if (fp && !buf.empty()) { fwrite(buf.data(), 1, buf.size(), fp); buf.clear(); } }
};

static inline char *put_u32(char *p, uint32_t v) {
    char tmp[10]; int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

static void write_line(clb_bed_writer *w, const std::string &name, const clb_interval &iv) {
    char line[64];
    char *p = line;
    *p++ = '\t'; p = put_u32(p, iv.start); *p++ = '\t'; p = put_u32(p, iv.end); *p++ = '\t';
    const char *sn = kStateName[iv.state < 6 ? iv.state : 0];
    const size_t sl = strlen(sn);
    memcpy(p, sn, sl); p += sl; *p++ = '\n';
    w->put(name.data(), name.size());
    w->put(line, (size_t)(p - line));
}

static inline bool binned_state(uint8_t s) { return s == CLB_CALLABLE || s == CLB_POOR_MAPPING_QUALITY || s == CLB_REF_N; }

extern "C" clb_bed_writer *clb_bed_writer_open(const char *path, uint32_t largest_contig_len) {
    clb_bed_writer *w = new clb_bed_writer();
    w->largest = largest_contig_len;
    if (!path) { w->in_memory = true; return w; }
    w->fp = fopen(path, "wb");                                // File::create truncates: callable_profiler.rs:31
    if (!w->fp) { delete w; return nullptr; }
    return w;
}

extern "C" int clb_bed_writer_add_contig(clb_bed_writer *w, const char *name, uint32_t contig_len, const clb_interval *iv, uint64_t n_iv,
                                         uint32_t *bins_inout, uint32_t n_bins, uint32_t stride, int *has_bins) {
    if (!w || !name) return CLB_E_INVALID;
    // intervals must tile [0, contig_len)
    uint32_t expect = 0;
    for (uint64_t i = 0; i < n_iv; i++) {
        if (iv[i].start != expect || iv[i].end <= iv[i].start || iv[i].state > 5) return CLB_E_INPUT;
        if (i > 0 && iv[i].state == iv[i - 1].state) return CLB_E_INPUT;
        expect = iv[i].end;
    }
    if (expect != contig_len) return CLB_E_INPUT;
    const std::string nm(name);
    bool any_range = false;
    // Q1: the first position of a contig (or finish_contig of an empty one) flushes the previous contig's
    // last run once more; Q2: that flush also pushes it into THIS contig's coverage_ranges.
    if (w->have_pending) {
        write_line(w, w->pend_name, w->pend);
        if (binned_state(w->pend.state)) {
            any_range = true;
            if (bins_inout && n_bins && stride) {
                uint32_t *row = bins_inout + (size_t)(w->pend.state == CLB_CALLABLE ? 0 : w->pend.state == CLB_POOR_MAPPING_QUALITY ? 1 : 2) * n_bins;
                // positions p in [start,end) with p / stride < n_bins  (histogram_plotter.rs:82-98)
                const uint64_t lim = (uint64_t)n_bins * stride;
                const uint64_t a = w->pend.start, b = std::min<uint64_t>(w->pend.end, lim);
                for (uint64_t p = a; p < b;) {
                    const uint64_t bi = p / stride, nxt = std::min<uint64_t>(b, (bi + 1) * stride);
                    row[bi] += (uint32_t)(nxt - p);
                    p = nxt;
                }
            }
        }
    }
    for (uint64_t i = 0; i < n_iv; i++) {
        write_line(w, nm, iv[i]);
        if (binned_state(iv[i].state)) any_range = true;
    }
    if (n_iv) { w->have_pending = true; w->pend_name = nm; w->pend = iv[n_iv - 1]; }
    if (has_bins) *has_bins = any_range ? 1 : 0;
    return CLB_OK;
}

extern "C" const char *clb_bed_writer_buffer(clb_bed_writer *w, uint64_t *len) {
    if (!w) return nullptr;
    if (len) *len = w->mem.size();
    return w->mem.data();
}

extern "C" int clb_bed_writer_close(clb_bed_writer *w) {
    if (!w) return CLB_E_INVALID;
    int rc = CLB_OK;
    if (w->fp) { w->flush(); if (fclose(w->fp) != 0) rc = CLB_E_IO; }
    delete w;
    return rc;
}
