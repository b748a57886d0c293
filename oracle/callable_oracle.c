/*
 * callable_oracle.c -- CPU ORACLE for the DecodingUsTools `coverage` / CallableLoci hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (decodingustools_b200/, the C-ABI
 * library, the CLI) may link, import or call this file.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker / CPU baseline.
 *
 * PARITY UNPINNED: the reference (Rust, rust-htslib 0.49.0 -> htslib C) cannot be built in
 * this image (no cargo/rustc, no htslib sources) and ships no tests or golden vectors for this
 * path (the only #[cfg(test)] in the tree is src/vg/framing.rs:114-165).  This file is a
 * restatement from code reading; it is pinned only by the hand-derived known-answer vectors
 * KA1-KA7 and by an independently written naive model
 * (oracle/naive_model.py).
 *
 * What is restated, single-threaded, in the reference's own order of operations:
 *   - htslib pileup engine as driven by rust-htslib (NOT in /root/reference; third-party,
 *     version unpinned because no Cargo.lock is checked in; call sites
 *     src/callable_loci/mod.rs:53-71): bam_plp_push admission with maxcnt, bam_plp64_next
 *     column generation, resolve_cigar2 per-column CIGAR cursor, bam_plp_auto driver.
 *   - process_position                     src/callable_loci/mod.rs:17-42
 *   - process_single_contig                src/callable_loci/mod.rs:44-147
 *   - CallableProfiler (classify/RLE/BED)  src/callable_loci/profilers/callable_profiler.rs:22-160
 *   - ContigProfiler (per-contig sums)     src/callable_loci/profilers/contig_profiler.rs:47-158
 *   - process_coverage_ranges + stride     src/callable_loci/utils/histogram_plotter.rs:74-102,412-441
 *   - build_coverage_export + natural sort src/callable_loci/report.rs:15-134,337-393
 *   - contig loop / largest_contig_length  src/api/coverage.rs:206-236
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* ------------------------------------------------------------------------------------------ */
/* Public structs (mirrored with ctypes in oracle/oracle.py)                                   */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    uint32_t min_depth;               /* options.rs:3  */
    uint32_t max_depth;               /* options.rs:4  */
    uint32_t min_depth_for_low_mapq;  /* options.rs:7  */
    uint8_t  min_mapping_quality;     /* options.rs:5  */
    uint8_t  min_base_quality;        /* options.rs:6  */
    uint8_t  max_low_mapq;            /* options.rs:8  */
    uint8_t  _pad;
    double   max_low_mapq_fraction;   /* options.rs:9  */
} orc_options;

/* One contig's records in BAM order (coordinate sorted, NOT admission filtered). */
typedef struct {
    uint64_t        n;
    const int32_t  *pos;        /* 0-based leftmost coordinate                */
    const uint16_t *flag;
    const uint8_t  *mapq;
    const uint32_t *cigar_off;  /* n+1 entries, index into cigar[]            */
    const uint32_t *cigar;      /* BAM encoding: len<<4 | op (MIDNSHP=X = 0-8) */
    const uint64_t *qual_off;   /* n+1 entries, byte offset into qual[]       */
    const uint8_t  *qual;       /* raw phred bytes                            */
    const uint32_t *name_id;    /* interned qname ids (mates share), or NULL  */
    uint32_t        n_names;    /* ids are < n_names                          */
} orc_reads;

typedef struct {
    uint64_t counts[6];         /* callable_profiler.rs:124-126, index = CalledState discriminant */
    uint64_t n_covered_bases;   /* contig_profiler.rs:79-82 */
    uint64_t summed_coverage;
    uint64_t summed_baseq;      /* contig_profiler.rs:69-70 */
    uint64_t summed_mapq;       /* contig_profiler.rs:74    */
    uint64_t quality_bases;     /* contig_profiler.rs:71    */
    uint64_t n_reads;           /* contig_profiler.rs:59-62 (u32 in the reference) */
    uint64_t n_admitted;        /* diagnostics: reads that entered the pileup */
    uint32_t length;
    uint32_t has_bins;          /* finish_contig only bins when coverage_ranges is non-empty */
    uint32_t n_bins;
    uint32_t stride;
} orc_contig_result;

typedef struct {
    double coverage_percent, average_depth, average_mapq, average_baseq, q30_percentage;
} orc_contig_floats;

typedef struct {
    uint64_t total_bases, callable_bases, total_unique_reads, contigs_analyzed;
    double   callable_percentage, average_depth;
    double   average_mapq, average_baseq, q30_percentage;
} orc_summary;

enum { ST_REF_N = 0, ST_CALLABLE = 1, ST_NO_COVERAGE = 2, ST_LOW_COVERAGE = 3,
       ST_EXCESSIVE_COVERAGE = 4, ST_POOR_MAPPING_QUALITY = 5 };   /* types.rs:36-43 */

static const char *STATE_NAME[6] = { "REF_N", "CALLABLE", "NO_COVERAGE", "LOW_COVERAGE",
                                     "EXCESSIVE_COVERAGE", "POOR_MAPPING_QUALITY" };

/* ------------------------------------------------------------------------------------------ */
/* Small growable buffers                                                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct { char *p; size_t n, cap; } sbuf;
static void sbuf_put(sbuf *b, const char *s, size_t n) {
    if (b->n + n + 1 > b->cap) {
        size_t nc = b->cap ? b->cap * 2 : 4096;
        while (nc < b->n + n + 1) nc *= 2;
        b->p = (char *)realloc(b->p, nc);
        b->cap = nc;
    }
    memcpy(b->p + b->n, s, n);
    b->n += n;
    b->p[b->n] = 0;
}

typedef struct { uint32_t start, end; int state; } cov_range;   /* types.rs:55-59 */
typedef struct { cov_range *p; size_t n, cap; } rvec;
static void rvec_push(rvec *v, cov_range r) {
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 256; v->p = (cov_range *)realloc(v->p, v->cap * sizeof(cov_range)); }
    v->p[v->n++] = r;
}

/* ------------------------------------------------------------------------------------------ */
/* The run object = one CallableProfiler + the per-contig ContigProfilers                      */
/* ------------------------------------------------------------------------------------------ */
#define ORC_MAX_NAME 256
typedef struct {
    char name[ORC_MAX_NAME];
    orc_contig_result res;
    uint32_t *bins[3];          /* callable, low_qual(POOR_MQ), ref_n   histogram_plotter.rs:76-78 */
} contig_slot;

typedef struct orc_run {
    orc_options opt;
    uint32_t largest_contig_length;      /* api/coverage.rs:210-215 */
    /* CallableProfiler state, callable_profiler.rs:11-19 */
    int      have_current;
    char     cur_contig[ORC_MAX_NAME];
    uint64_t cur_start, cur_end;
    int      cur_state;
    sbuf     bed;
    rvec     ranges;                     /* coverage_ranges */
    contig_slot *contigs; size_t n_contigs, cap_contigs;
    contig_slot *active;                 /* contig whose counts process_state updates */
    char err[256];
} orc_run;

orc_run *orc_run_new(const orc_options *opt, uint32_t largest_contig_length) {
    orc_run *r = (orc_run *)calloc(1, sizeof(orc_run));
    r->opt = *opt;
    r->largest_contig_length = largest_contig_length;
    return r;
}

void orc_run_free(orc_run *r) {
    if (!r) return;
    for (size_t i = 0; i < r->n_contigs; i++) for (int k = 0; k < 3; k++) free(r->contigs[i].bins[k]);
    free(r->contigs); free(r->bed.p); free(r->ranges.p); free(r);
}

const char *orc_bed(orc_run *r, uint64_t *len) { *len = r->bed.n; return r->bed.p ? r->bed.p : ""; }
const char *orc_last_error(orc_run *r) { return r->err; }

/* callable_profiler.rs:39-62 */
static void write_state(orc_run *r) {
    if (!r->have_current) return;
    char line[ORC_MAX_NAME + 96];
    int n = snprintf(line, sizeof line, "%s\t%llu\t%llu\t%s\n", r->cur_contig,
                     (unsigned long long)r->cur_start, (unsigned long long)r->cur_end, STATE_NAME[r->cur_state]);
    sbuf_put(&r->bed, line, (size_t)n);
    if (r->cur_state == ST_CALLABLE || r->cur_state == ST_POOR_MAPPING_QUALITY || r->cur_state == ST_REF_N) {
        cov_range cr = { (uint32_t)r->cur_start, (uint32_t)r->cur_end, r->cur_state };
        rvec_push(&r->ranges, cr);
    }
}

/* callable_profiler.rs:122-155 */
static void process_state(orc_run *r, const char *contig, uint64_t pos, int state) {
    r->active->res.counts[state] += 1;                   /* :124-126 (keyed by contig name) */
    if (!r->have_current) {                              /* :128-141 */
        if (state == ST_REF_N) {
            strncpy(r->cur_contig, contig, ORC_MAX_NAME - 1);
            r->cur_start = 0; r->cur_end = pos + 1; r->cur_state = state; r->have_current = 1;
        } else {
            if (pos > 0) {
                strncpy(r->cur_contig, contig, ORC_MAX_NAME - 1);
                r->cur_start = 0; r->cur_end = pos; r->cur_state = ST_REF_N; r->have_current = 1;
                write_state(r);
            }
            strncpy(r->cur_contig, contig, ORC_MAX_NAME - 1);
            r->cur_start = pos; r->cur_end = pos + 1; r->cur_state = state; r->have_current = 1;
        }
        return;
    }
    if (strcmp(r->cur_contig, contig) == 0 && r->cur_state == state) {
        r->cur_end = pos + 1;                            /* :145-146 */
    } else {
        write_state(r);                                  /* :148-150 */
        strncpy(r->cur_contig, contig, ORC_MAX_NAME - 1);
        r->cur_start = pos; r->cur_end = pos + 1; r->cur_state = state;
    }
}

/* callable_profiler.rs:89-120 */
static int classify(const orc_options *o, uint8_t ref_base, uint32_t raw, uint32_t qc, uint32_t low) {
    int is_low_mapq = raw >= o->min_depth_for_low_mapq &&
                      ((double)low / (double)raw) > o->max_low_mapq_fraction;
    if (ref_base == 'N' || ref_base == 'n') return ST_REF_N;
    if (raw == 0) return ST_NO_COVERAGE;
    if (is_low_mapq) return ST_POOR_MAPPING_QUALITY;
    if (qc < o->min_depth) return ST_LOW_COVERAGE;
    if (o->max_depth > 0 && qc > o->max_depth) return ST_EXCESSIVE_COVERAGE;
    return ST_CALLABLE;
}

/* histogram_plotter.rs:74-102 and 412-441; called from finish_contig callable_profiler.rs:64-87 */
static int finish_contig(orc_run *r, contig_slot *c) {
    write_state(r);                                      /* :65, current_state is NOT cleared (quirk Q1) */
    if (r->ranges.n > 0) {                               /* :67 */
        uint32_t largest = strcmp(c->name, "chrM") == 0 ? c->res.length : r->largest_contig_length;
        uint32_t stride = strcmp(c->name, "chrM") == 0 ? (16569u + 200u - 1u) / 200u
                                                       : (uint32_t)(((uint64_t)largest + 2000u - 1u) / 2000u);
        if (stride == 0) { snprintf(r->err, sizeof r->err, "stride 0 (reference would panic: division by zero)"); return -1; }
        uint32_t n_bins = c->res.length / stride + 1;    /* :75 with min_cutoff 0, max_cutoff contig_length */
        for (int k = 0; k < 3; k++) c->bins[k] = (uint32_t *)calloc(n_bins, sizeof(uint32_t));
        for (size_t i = 0; i < r->ranges.n; i++) {
            cov_range cr = r->ranges.p[i];
            int k = cr.state == ST_CALLABLE ? 0 : cr.state == ST_POOR_MAPPING_QUALITY ? 1 : 2;
            for (uint32_t p = cr.start; p < cr.end; p++) {   /* :82-98, per position on purpose */
                uint32_t idx = p / stride;
                if (idx < n_bins) c->bins[k][idx] += 1;
            }
        }
        c->res.has_bins = 1; c->res.n_bins = n_bins; c->res.stride = stride;
        r->ranges.n = 0;                                 /* std::mem::take */
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* htslib-style pileup iterator (bam_plp_*), restated.                                         */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int64_t  beg, end;     /* lbnode_t beg/end: end = pos + cigar2rlen (raw, may equal beg) */
    uint64_t r;            /* record index */
    /* resolve_cigar2 cursor: op index k, reference coordinate x and query offset y at op start */
    uint32_t k; int64_t x; int64_t y;
} plp_node;

typedef struct {
    const orc_reads *rd;
    int64_t  tid;              /* tid of the contig we were fetched on (all records share it) */
    /* iterator state */
    int64_t  it_tid, it_pos;   /* iter->tid, iter->pos (start 0,0 from calloc)   */
    int64_t  max_tid, max_pos; /* iter->max_tid, iter->max_pos (start -1,-1)     */
    int      is_eof;
    uint64_t next_rec;         /* next record the read callback would return     */
    uint32_t maxcnt;           /* bam_plp_set_maxcnt                             */
    uint64_t cnt;              /* mempool live-node count INCLUDING the tail placeholder */
    plp_node *list; size_t n, cap;     /* head..tail-1 in push order                    */
    uint64_t n_admitted;
    int      error;
} plp_iter;

static int op_consumes_ref(uint32_t op)   { return op == 0 || op == 2 || op == 3 || op == 7 || op == 8; }
static int op_consumes_query(uint32_t op) { return op == 0 || op == 1 || op == 4 || op == 7 || op == 8; }

static void plp_init(plp_iter *it, const orc_reads *rd, int64_t tid, uint32_t maxcnt) {
    memset(it, 0, sizeof *it);
    it->rd = rd; it->tid = tid; it->maxcnt = maxcnt;
    it->max_tid = -1; it->max_pos = -1;
    it->cnt = 1;               /* bam_plp_init allocates the tail placeholder */
}

/* bam_plp_push for one real record */
static void plp_push(plp_iter *it, uint64_t r) {
    const orc_reads *rd = it->rd;
    int64_t pos = rd->pos[r];
    if (rd->flag[r] & 0x4) return;                              /* BAM_FUNMAP only */
    if (it->it_tid == it->tid && it->it_pos == pos && it->cnt > it->maxcnt) return;   /* depth cap */
    int64_t rlen = 0;
    for (uint32_t c = rd->cigar_off[r]; c < rd->cigar_off[r + 1]; c++)
        if (op_consumes_ref(rd->cigar[c] & 0xf)) rlen += rd->cigar[c] >> 4;
    int64_t end = pos + rlen;
    if (it->tid < it->max_tid || (it->tid == it->max_tid && pos < it->max_pos)) { it->error = 1; return; }  /* unsorted */
    it->max_tid = it->tid; it->max_pos = pos;
    if (end > it->it_pos || it->tid > it->it_tid) {             /* keep node, allocate new tail */
        if (it->n == it->cap) { it->cap = it->cap ? it->cap * 2 : 1024; it->list = (plp_node *)realloc(it->list, it->cap * sizeof(plp_node)); }
        plp_node *nd = &it->list[it->n++];
        nd->beg = pos; nd->end = end; nd->r = r; nd->k = rd->cigar_off[r]; nd->x = pos; nd->y = 0;
        it->cnt++;
        it->n_admitted++;
    }
}

typedef struct { uint64_t r; int is_del; int is_refskip; int64_t qpos; } plp_aln;

/* One pass of the column loop of bam_plp64_next.  Returns 1 with a non-empty column in out[],
 * 0 when no column can be produced yet (need more reads) or at end. */
static int plp_next(plp_iter *it, int64_t *o_tid, int64_t *o_pos, plp_aln **out, size_t *out_cap, size_t *o_n) {
    const orc_reads *rd = it->rd;
    if (it->is_eof && it->n == 0) return 0;
    while (it->is_eof || it->max_tid > it->it_tid || (it->max_tid == it->it_tid && it->max_pos > it->it_pos)) {
        size_t n_plp = 0, w = 0;
        for (size_t i = 0; i < it->n; i++) {
            plp_node *p = &it->list[i];
            if (it->tid < it->it_tid || (it->tid == it->it_tid && p->end <= it->it_pos)) {
                it->cnt--;                                  /* mp_free */
                continue;                                   /* unlink */
            }
            if (it->tid == it->it_tid && p->beg <= it->it_pos) {
                /* resolve_cigar2: find the reference-consuming op that contains it_pos */
                for (;;) {
                    uint32_t op = rd->cigar[p->k] & 0xf, len = rd->cigar[p->k] >> 4;
                    if (op_consumes_ref(op) && it->it_pos < p->x + (int64_t)len) break;
                    if (op_consumes_ref(op))   p->x += len;
                    if (op_consumes_query(op)) p->y += len;
                    p->k++;
                }
                uint32_t op = rd->cigar[p->k] & 0xf;
                if (n_plp == *out_cap) { *out_cap = *out_cap ? *out_cap * 2 : 1024; *out = (plp_aln *)realloc(*out, *out_cap * sizeof(plp_aln)); }
                plp_aln *a = &(*out)[n_plp++];
                a->r = p->r;
                a->is_del = (op == 2 || op == 3);
                a->is_refskip = (op == 3);
                a->qpos = p->y + (it->it_pos - p->x);       /* only meaningful for M/=/X */
            }
            if (w != i) it->list[w] = *p;
            w++;
        }
        it->n = w;
        *o_n = n_plp; *o_tid = it->it_tid; *o_pos = it->it_pos;
        if (it->n > 0) {
            if (it->it_tid < it->tid) { it->it_tid = it->tid; it->it_pos = it->list[0].beg; }
            else if (it->it_pos < it->list[0].beg) it->it_pos = it->list[0].beg;
            else it->it_pos++;
        } else {
            /* head == tail placeholder: htslib reads the placeholder's stale fields here; the only
             * way to get here with the loop still running is is_eof, where we stop right below. */
            it->it_pos++;
        }
        if (n_plp) return 1;
        if (it->is_eof && it->n == 0) break;
    }
    return 0;
}

/* bam_plp_auto: produce the next non-empty column, reading records as needed. */
static int plp_auto(plp_iter *it, int64_t *o_tid, int64_t *o_pos, plp_aln **out, size_t *out_cap, size_t *o_n) {
    if (plp_next(it, o_tid, o_pos, out, out_cap, o_n)) return 1;
    if (it->is_eof) return 0;
    while (it->next_rec < it->rd->n) {
        plp_push(it, it->next_rec++);
        if (it->error) return 0;
        if (plp_next(it, o_tid, o_pos, out, out_cap, o_n)) return 1;
    }
    it->is_eof = 1;                                          /* bam_plp_push(iter, 0) */
    return plp_next(it, o_tid, o_pos, out, out_cap, o_n);
}

/* Stand-alone admission (A0) so tests can pin the keep mask: keep[r] = 1 iff the record entered
 * the pileup list.  Runs the very same iterator. */
int orc_admit(const orc_reads *rd, int32_t tid, uint32_t maxcnt, uint8_t *keep) {
    plp_iter it; plp_init(&it, rd, tid, maxcnt);
    plp_aln *col = NULL; size_t cap = 0, n = 0; int64_t t, p;
    memset(keep, 0, rd->n);
    /* replicate plp_auto but record admissions */
    for (;;) {
        while (plp_next(&it, &t, &p, &col, &cap, &n)) {}
        if (it.next_rec >= rd->n) break;
        uint64_t before = it.n_admitted;
        plp_push(&it, it.next_rec);
        if (it.error) { free(col); free(it.list); return -1; }
        if (it.n_admitted != before) keep[it.next_rec] = 1;
        it.next_rec++;
    }
    free(col); free(it.list);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* process_single_contig (mod.rs:44-147)                                                       */
/* ------------------------------------------------------------------------------------------ */
static uint8_t fetch_ref(const uint8_t *ref, uint64_t ref_len, uint64_t p) {
    /* fasta.fetch_seq(contig,p,p).first().unwrap_or(b'N')  mod.rs:79-80 */
    return (ref && p < ref_len) ? ref[p] : (uint8_t)'N';
}

static contig_slot *new_contig(orc_run *r, const char *name, uint32_t len) {
    if (r->n_contigs == r->cap_contigs) {
        r->cap_contigs = r->cap_contigs ? r->cap_contigs * 2 : 32;
        r->contigs = (contig_slot *)realloc(r->contigs, r->cap_contigs * sizeof(contig_slot));
    }
    contig_slot *c = &r->contigs[r->n_contigs++];
    memset(c, 0, sizeof *c);
    strncpy(c->name, name, ORC_MAX_NAME - 1);
    c->res.length = len;
    return c;
}

/* Process the next contig (callers pass contigs in ascending tid, api/coverage.rs:229-235).
 * dbg_* (each contig_len entries) may be NULL.  Returns 0 or -1 (see orc_last_error). */
int orc_process_contig(orc_run *r, const char *name, int32_t tid, uint32_t contig_len,
                       const uint8_t *ref, uint64_t ref_len, const orc_reads *rd,
                       orc_contig_result *out,
                       uint32_t *dbg_raw, uint32_t *dbg_qc, uint32_t *dbg_low, uint8_t *dbg_state) {
    const orc_options *o = &r->opt;
    contig_slot *c = new_contig(r, name, contig_len);
    r->active = c;
    uint8_t *seen = (rd->name_id && rd->n_names) ? (uint8_t *)calloc(rd->n_names, 1) : NULL;

    plp_iter it;
    plp_init(&it, rd, tid, o->max_depth > 0 ? o->max_depth : 500);      /* mod.rs:56-60 */
    plp_aln *col = NULL; size_t cap = 0, n = 0; int64_t c_tid = 0, c_pos = 0;
    uint32_t current_pos = 0;

    while (plp_auto(&it, &c_tid, &c_pos, &col, &cap, &n)) {
        if (c_tid != tid) break;                                         /* mod.rs:67-69 */
        uint32_t pos = (uint32_t)c_pos;
        while (current_pos < pos) {                                      /* gap fill mod.rs:74-93 */
            int st = classify(o, fetch_ref(ref, ref_len, current_pos), 0, 0, 0);
            if (dbg_state && current_pos < contig_len) { dbg_state[current_pos] = (uint8_t)st; }
            process_state(r, name, current_pos, st);
            current_pos++;
        }
        /* process_position mod.rs:17-42 */
        uint32_t raw = 0, qc = 0, low = 0;
        for (size_t i = 0; i < n; i++) {
            uint64_t rr = col[i].r;
            raw++;
            uint8_t mq = rd->mapq[rr];
            if (mq <= o->max_low_mapq) low++;
            if (mq >= o->min_mapping_quality) {
                if (!col[i].is_del) {                                     /* Alignment::qpos() is Some */
                    uint64_t lq = rd->qual_off[rr + 1] - rd->qual_off[rr];
                    if ((uint64_t)col[i].qpos < lq) {                     /* record.qual().get(qpos) */
                        uint8_t q = rd->qual[rd->qual_off[rr] + (uint64_t)col[i].qpos];
                        if (q >= o->min_base_quality /* || is_del: dead, qpos is None for deletions */) qc++;
                    }
                }
            }
        }
        int st = classify(o, fetch_ref(ref, ref_len, pos), raw, qc, low);
        if (pos < contig_len) {
            if (dbg_raw) dbg_raw[pos] = raw;
            if (dbg_qc) dbg_qc[pos] = qc;
            if (dbg_low) dbg_low[pos] = low;
            if (dbg_state) dbg_state[pos] = (uint8_t)st;
        }
        process_state(r, name, pos, st);
        /* ContigProfiler::process_position contig_profiler.rs:47-83 (second pass over the column) */
        for (size_t i = 0; i < n; i++) {
            uint64_t rr = col[i].r;
            uint8_t mq = rd->mapq[rr];
            if (seen) { uint32_t id = rd->name_id[rr]; if (!seen[id]) { seen[id] = 1; c->res.n_reads++; } }
            if (mq >= o->min_mapping_quality) {
                if (!col[i].is_del) {
                    uint64_t lq = rd->qual_off[rr + 1] - rd->qual_off[rr];
                    if ((uint64_t)col[i].qpos < lq) {
                        uint8_t q = rd->qual[rd->qual_off[rr] + (uint64_t)col[i].qpos];
                        if (q >= o->min_base_quality) { c->res.summed_baseq += q; c->res.quality_bases += 1; }
                    }
                }
                c->res.summed_mapq += mq;
            }
        }
        if (raw > 0) { c->res.n_covered_bases += 1; c->res.summed_coverage += raw; }
        current_pos = pos + 1;
    }
    if (it.error) { snprintf(r->err, sizeof r->err, "records are not coordinate sorted"); free(col); free(it.list); free(seen); return -1; }
    while (current_pos < contig_len) {                                   /* tail fill mod.rs:123-142 */
        int st = classify(o, fetch_ref(ref, ref_len, current_pos), 0, 0, 0);
        if (dbg_state) dbg_state[current_pos] = (uint8_t)st;
        process_state(r, name, current_pos, st);
        current_pos++;
    }
    c->res.n_admitted = it.n_admitted;
    free(col); free(it.list); free(seen);
    if (finish_contig(r, c) != 0) return -1;                             /* mod.rs:144-145 */
    if (out) *out = c->res;
    return 0;
}

/* Copy the three bin arrays of contig index i (processing order). */
int orc_get_bins(orc_run *r, uint32_t i, uint32_t *callable, uint32_t *low_qual, uint32_t *ref_n) {
    if (i >= r->n_contigs || !r->contigs[i].res.has_bins) return -1;
    uint32_t n = r->contigs[i].res.n_bins;
    memcpy(callable, r->contigs[i].bins[0], n * 4);
    memcpy(low_qual, r->contigs[i].bins[1], n * 4);
    memcpy(ref_n, r->contigs[i].bins[2], n * 4);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Final floats: contig_profiler.rs:123-158, report.rs:15-134, natural order report.rs:337-393 */
/* ------------------------------------------------------------------------------------------ */
static size_t split_pos(const char *s) {      /* report.rs:386-393 */
    size_t i = 0;
    for (; s[i]; i++) if ((s[i] >= '0' && s[i] <= '9') || s[i] == 'X' || s[i] == 'Y' || s[i] == 'M') break;
    return i;
}
static void order_key(const char *suf, unsigned *cat, uint32_t *num) {   /* report.rs:353-366 */
    /* Rust str::parse::<u32>: optional leading '+', digits only, no overflow */
    const char *p = suf; if (*p == '+') p++;
    int ok = (*p != 0); uint64_t v = 0;
    for (const char *q = p; *q; q++) { if (*q < '0' || *q > '9') { ok = 0; break; } v = v * 10 + (uint64_t)(*q - '0'); if (v > 0xffffffffull) { ok = 0; break; } }
    if (ok) { *cat = 0; *num = (uint32_t)v; return; }
    *num = 0;
    if (!strcmp(suf, "X")) *cat = 1; else if (!strcmp(suf, "Y")) *cat = 2;
    else if (!strcmp(suf, "M") || !strcmp(suf, "MT")) *cat = 3; else *cat = 4;
}
int orc_compare_contig_names(const char *a, const char *b) {           /* report.rs:339-383 */
    size_t sa = split_pos(a), sb = split_pos(b);
    size_t m = sa < sb ? sa : sb;
    int c = memcmp(a, b, m);
    if (c == 0 && sa != sb) c = sa < sb ? -1 : 1;
    if (c) return c < 0 ? -1 : 1;
    unsigned ca, cb; uint32_t na, nb;
    order_key(a + sa, &ca, &na); order_key(b + sb, &cb, &nb);
    if (ca != cb) return ca < cb ? -1 : 1;
    if (ca == 0) return na < nb ? -1 : na > nb ? 1 : 0;
    c = strcmp(a + sa, b + sb);
    return c < 0 ? -1 : c > 0 ? 1 : 0;
}

static void contig_floats(const orc_contig_result *s, orc_contig_floats *f) {
    f->coverage_percent = s->length > 0 ? ((double)s->n_covered_bases / (double)s->length) * 100.0 : 0.0;
    f->average_depth = s->n_covered_bases > 0 ? (double)s->summed_coverage / (double)s->n_covered_bases : 0.0;
    f->average_mapq = s->quality_bases > 0 ? (double)s->summed_mapq / (double)s->quality_bases : 0.0;
    f->average_baseq = s->quality_bases > 0 ? (double)s->summed_baseq / (double)s->quality_bases : 0.0;
    if (s->quality_bases > 0) {
        if (f->average_baseq >= 30.0) f->q30_percentage = 100.0;
        else if (f->average_baseq < 20.0) f->q30_percentage = 0.0;
        else f->q30_percentage = ((f->average_baseq - 20.0) / 10.0) * 100.0;
    } else f->q30_percentage = 0.0;
}

/* order[] receives contig indices (processing order) sorted the way the report sorts them;
 * floats[] is indexed like order[] (i.e. natural order). */
int orc_build_export(orc_run *r, uint32_t *order, orc_contig_floats *floats, orc_summary *sum) {
    size_t n = r->n_contigs;
    for (size_t i = 0; i < n; i++) order[i] = (uint32_t)i;
    for (size_t i = 1; i < n; i++) {            /* stable insertion sort == Rust's stable sort_by */
        uint32_t v = order[i]; size_t j = i;
        while (j > 0 && orc_compare_contig_names(r->contigs[order[j - 1]].name, r->contigs[v].name) > 0) { order[j] = order[j - 1]; j--; }
        order[j] = v;
    }
    uint64_t total_bases = 0, callable_bases = 0, q30_bases = 0, total_qpos = 0, total_unique = 0;
    double total_depth = 0, total_mapq = 0, total_baseq = 0;
    for (size_t i = 0; i < n; i++) {
        const orc_contig_result *s = &r->contigs[order[i]].res;
        contig_floats(s, &floats[i]);
        total_bases += s->length;
        callable_bases += s->counts[ST_CALLABLE];
        total_depth += floats[i].average_depth * (double)s->length;
        total_mapq += floats[i].average_mapq * (double)s->length;
        total_baseq += floats[i].average_baseq * (double)s->length;
        q30_bases += (uint64_t)(floats[i].q30_percentage / 100.0 * (double)s->length);
        total_qpos += s->length;
        total_unique += (uint32_t)s->n_reads;   /* n_reads is u32 in the reference */
    }
    memset(sum, 0, sizeof *sum);
    sum->total_bases = total_bases; sum->callable_bases = callable_bases;
    sum->total_unique_reads = total_unique; sum->contigs_analyzed = n;
    sum->average_depth = total_bases > 0 ? total_depth / (double)total_bases : 0.0;
    sum->callable_percentage = total_bases > 0 ? ((double)callable_bases / (double)total_bases) * 100.0 : 0.0;
    sum->average_mapq = total_qpos > 0 ? total_mapq / (double)total_qpos : 0.0;
    sum->average_baseq = total_qpos > 0 ? total_baseq / (double)total_qpos : 0.0;
    sum->q30_percentage = total_qpos > 0 ? ((double)q30_bases / (double)total_qpos) * 100.0 : 0.0;
    return 0;
}

uint32_t orc_n_contigs(orc_run *r) { return (uint32_t)r->n_contigs; }
const char *orc_contig_name(orc_run *r, uint32_t i) { return i < r->n_contigs ? r->contigs[i].name : ""; }
int orc_contig_result_get(orc_run *r, uint32_t i, orc_contig_result *out) {
    if (i >= r->n_contigs) return -1;
    *out = r->contigs[i].res; return 0;
}
