// clb_kernels.cuh -- hand-written sm_100a kernels of the CallableLoci hot path.
//
// Replaces the reference's per-base CPU loop:
//   htslib bam_plp64_next + process_position          /root/reference/src/callable_loci/mod.rs:17-42,65-120
//   CallableProfiler::process_position/process_state  .../profilers/callable_profiler.rs:89-155
//   ContigProfiler::process_position                  .../profilers/contig_profiler.rs:47-83
//   HistogramPlotter::process_coverage_ranges         .../utils/histogram_plotter.rs:74-102
//
// Design: one CTA owns a reference window of WREAL positions plus one halo position
// to its left.  Warps autonomously pull 32 candidate reads at a time (coalesced column loads), walk the
// CIGARs (lane-serial for short CIGARs, warp-cooperative prefix sums for long ones), and turn every
// read / M-segment into two shared-memory difference-array updates instead of one update per base.
// Base qualities are streamed once with 16-byte loads; only bases that FAIL the quality threshold touch
// a per-position counter.  After a block scan the window is classified, run boundaries are compacted
// into interval records, and counters / bins are reduced per CTA before a handful of global atomics.
// No per-base array ever reaches HBM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace clb {

constexpr int NT = 256;               // threads per CTA
constexpr int NWARPS = NT / 32;
constexpr int PPT = 16;               // window entries per thread in the classify phase
constexpr int WN = NT * PPT;          // entries per window; entry 0 is the halo position (window start - 1)
constexpr int WREAL = WN - 1;         // reference positions owned by one window
constexpr int MAXSEG = 2;             // M-like segments a "simple" read may contribute
constexpr int FAST_OPS = 6;           // CIGAR ops walked lane-serially; longer CIGARs go warp-cooperative
constexpr int CHUNK_CAP = 352;        // 16-byte quality chunks mapped per warp round
constexpr int NFIRST = 128;             // low-MAPQ threshold table entries cached in shared memory
constexpr unsigned FULL = 0xffffffffu;

constexpr int A_WORDS = ((WN + WN / 16 + 4) / 4) * 4;              // padded u32 difference arrays
constexpr int LQ_WORDS = (((WN + 2 * (WN / 16)) / 2 + 4) / 4) * 4; // padded packed-u16 low-BQ counters
constexpr int STAT_STRIDE = 16;       // one 128-byte line per global counter (u64 units)

enum { ST_REF_N = 0, ST_CALLABLE = 1, ST_NO_COVERAGE = 2, ST_LOW_COVERAGE = 3, ST_EXCESSIVE = 4, ST_POOR_MAPQ = 5 };
enum { S_COUNT0 = 0, S_COVERED = 6, S_SUMCOV = 7, S_SUMBQ = 8, S_SUMMAPQ = 9, S_QBASES = 10, S_QBASES_B = 11, N_STATS = 12 };
enum { ERR_QUAL_SPAN = 1, ERR_REC_OVERFLOW = 2, ERR_DEPTH = 4, ERR_UNSORTED = 8, ERR_OFFSETS = 16 };

struct KParams {
    // packed read columns (device)
    const int32_t  *pos;
    const uint16_t *flag;
    const uint8_t  *mapq;
    const uint32_t *cigar_off;
    const uint32_t *cigar;
    const uint64_t *qual_off;
    const uint8_t  *qual;
    const uint32_t *read_end;      // optional (long-read mode): pos + reference span, else nullptr
    // contig / region
    const uint32_t *nmask;         // bit-packed REF_N mask, zero padded past the contig end
    uint32_t region_start, region_end;
    // options
    uint32_t min_depth, max_depth, min_depth_for_low_mapq;
    uint32_t min_mapq, min_bq, max_low_mapq;
    const uint32_t *first_tab;     // [65536] smallest low count with low/raw > fraction (f64, exact)
    // windows
    const uint32_t *win_rlo, *win_rhi;
    uint32_t win_first;
    // outputs
    unsigned long long *stats;     // [N_STATS * STAT_STRIDE]
    unsigned long long *bins;      // [3][n_bins]
    uint32_t n_bins, stride;
    unsigned long long *rec;       // boundary records: pos | state << 32 | soft << 40
    uint32_t rec_cap;
    uint32_t *rec_cursor;
    uint2 *win_tab;                // per window: (first record, record count)
    uint32_t *err;
    // optional per-base debug output, indexed by position - region_start
    uint32_t *dbg_raw, *dbg_qc, *dbg_low;
    uint8_t *dbg_state;
};

// padded shared-memory indices: 16 consecutive entries per thread -> lane stride 17 words (9 for u16)
__device__ __forceinline__ uint32_t pidx(uint32_t e) { return e + (e >> 4); }
__device__ __forceinline__ uint32_t pidx16(uint32_t e) { return e + ((e >> 4) << 1); }

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ bool op_is_m(uint32_t op) { return op == 0 || op == 7 || op == 8; }
__device__ __forceinline__ bool op_ref(uint32_t op) { return op == 0 || op == 2 || op == 3 || op == 7 || op == 8; }
__device__ __forceinline__ bool op_qry(uint32_t op) { return op == 0 || op == 1 || op == 4 || op == 7 || op == 8; }

struct Seg { uint32_t qrel, rrel, len; };

// Per-CTA constants + shared-memory views
struct Win {
    long long wb, wend;            // position of entry 0, exclusive end of positions handled
    uint64_t qbase;                // 16-byte aligned byte offset of the window's first candidate quality
    const uint8_t *qual;
    uint32_t *sA, *sB, *sLQ;
    uint32_t min_bq, min_mapq, max_low_mapq;
};

// SWAR: 0x80 in every byte of x that is < T (unsigned), for any T in 0..255.
__device__ __forceinline__ uint32_t bytes_lt(uint32_t x, uint32_t t_low, uint32_t t_hi) {
    const uint32_t H = 0x80808080u;
    uint32_t d = (x | H) - t_low;                        // bit7 = ((x & 0x7f) >= (T & 0x7f)), no cross-byte borrow
    return ((~x & t_hi) | (~(x ^ t_hi) & ~d)) & H;
}

// 0xFF in every byte whose bit 7 is set (PRMT sign-replicate mode; __byte_perm() ignores the replicate bit).
__device__ __forceinline__ uint32_t bytes_from_msb(uint32_t x) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(r) : "r"(x));
    return r;
}

// Difference-array update for one M-like segment (reads with mapq >= min_mapq only); returns the part of
// the segment inside the window proper (entries >= 1) as a Seg.  rp/qp: reference/query coordinate of
// the op start; q0: absolute byte offset of the read's qualities; lq: its quality length.
__device__ __forceinline__ bool emit_m(const Win &W, long long rp, uint32_t qp, uint32_t len, uint64_t q0, uint32_t lq, Seg &out) {
    if (qp >= lq) return false;                          // record.qual().get(qpos) == None
    len = min(len, lq - qp);
    long long s = max(rp, W.wb), e = min(rp + (long long)len, W.wend);
    if (e <= s) return false;
    uint32_t e0 = (uint32_t)(s - W.wb), e1 = (uint32_t)(e - W.wb);
    atomicAdd(&W.sB[pidx(e0)], 1u);
    if (e1 < (uint32_t)WN) atomicAdd(&W.sB[pidx(e1)], 0xffffffffu);
    if (e0 == 0) {                                       // covers the halo position: test its one base here
        uint8_t q = W.qual[q0 + qp + (uint64_t)(W.wb - rp)];
        if (q < W.min_bq) atomicAdd(&W.sLQ[0], 1u);
        e0 = 1;
        if (e1 <= 1) return false;
    }
    out.rrel = e0;
    out.len = e1 - e0;
    out.qrel = (uint32_t)(q0 - W.qbase) + qp + (uint32_t)((W.wb + (long long)e0) - rp);
    return true;
}

// Difference-array update for a whole read (raw depth + low-MAPQ depth packed as lo16|hi16).
__device__ __forceinline__ void emit_read(const Win &W, long long p, long long end, uint32_t mq, unsigned long long &acc_mapq) {
    long long s = max(p, W.wb), e = min(end, W.wend);
    if (e <= s) return;
    uint32_t e0 = (uint32_t)(s - W.wb), e1 = (uint32_t)(e - W.wb);
    uint32_t delta = 1u + ((mq <= W.max_low_mapq) ? 0x10000u : 0u);
    atomicAdd(&W.sA[pidx(e0)], delta);
    if (e1 < (uint32_t)WN) atomicAdd(&W.sA[pidx(e1)], 0u - delta);
    uint32_t rs = max(e0, 1u);
    if (mq >= W.min_mapq && e1 > rs) acc_mapq += (unsigned long long)mq * (e1 - rs);
}

// One 16-byte quality chunk of one segment.
__device__ __forceinline__ void process_chunk(const Win &W, uint4 d, uint32_t f, uint4 v, const uint4 *sMaskLo, const uint4 *sMaskHi,
                                              uint32_t t_low, uint32_t t_hi, uint32_t &acc_sum, uint32_t &acc_cnt) {
    const uint32_t c = f - d.w;
    const uint32_t head = d.x & 15u;
    const uint32_t lo = c == 0 ? head : 0u;
    const uint32_t rem = head + d.z - 16u * c;            // bytes from chunk start to segment end (>= 1)
    const uint32_t hi = min(16u, rem);
    const uint4 ml = sMaskLo[lo], mh = sMaskHi[hi];       // 0xFF in bytes outside [lo, hi)
    v.x |= ml.x | mh.x; v.y |= ml.y | mh.y; v.z |= ml.z | mh.z; v.w |= ml.w | mh.w;
    const uint32_t l0 = bytes_lt(v.x, t_low, t_hi), l1 = bytes_lt(v.y, t_low, t_hi);
    const uint32_t l2 = bytes_lt(v.z, t_low, t_hi), l3 = bytes_lt(v.w, t_low, t_hi);
    uint32_t sum = __dp4a(v.x, 0x01010101u, __dp4a(v.y, 0x01010101u, __dp4a(v.z, 0x01010101u, __dp4a(v.w, 0x01010101u, 0u))));
    const uint32_t ninv = lo + (16u - hi);
    sum -= 255u * ninv;
    uint32_t cnt = 16u - ninv;
    if (l0 | l1 | l2 | l3) {
        // bit (8*j + w) set <=> byte j of word w fails, i.e. chunk byte 4*w + j
        uint32_t m = (l0 >> 7) | (l1 >> 6) | (l2 >> 5) | (l3 >> 4);
        cnt -= __popc(m);
        sum -= __dp4a(v.x & bytes_from_msb(l0), 0x01010101u, 0u) + __dp4a(v.y & bytes_from_msb(l1), 0x01010101u, 0u)
             + __dp4a(v.z & bytes_from_msb(l2), 0x01010101u, 0u) + __dp4a(v.w & bytes_from_msb(l3), 0x01010101u, 0u);
        const uint32_t e_chunk = d.y + 16u * c - head;    // entry of chunk byte 0 (may "underflow" for c == 0; fixed by + byte)
        while (m) {
            const uint32_t b = __ffs(m) - 1; m &= m - 1;
            const uint32_t e = e_chunk + ((b & 7u) << 2) + (b >> 3);
            const uint32_t i16 = pidx16(e);
            atomicAdd(&W.sLQ[i16 >> 1], 1u << ((i16 & 1u) << 4));
        }
    }
    acc_sum += sum; acc_cnt += cnt;
}

// Warp-collective: stream the qualities of the segments held in the lanes' registers.
__device__ __forceinline__ void process_segments(const Win &W, int nseg, const Seg (&seg)[MAXSEG], uint4 *myDesc, uint8_t *myMap,
                                                 const uint4 *sMaskLo, const uint4 *sMaskHi, uint32_t t_low, uint32_t t_hi,
                                                 uint32_t &acc_sum, uint32_t &acc_cnt, int lane) {
    uint32_t nc[MAXSEG]; uint32_t nch = 0;
#pragma unroll
    for (int k = 0; k < MAXSEG; k++) { nc[k] = k < nseg ? (((seg[k].qrel & 15u) + seg[k].len + 15u) >> 4) : 0u; nch += nc[k]; }
    uint32_t incl = nch;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) { uint32_t t = __shfl_up_sync(FULL, incl, dd); if (lane >= dd) incl += t; }
    const uint32_t total = __shfl_sync(FULL, incl, 31);
    if (total == 0) return;
    const uint32_t cf = incl - nch;
    {
        uint32_t cfk = cf;
#pragma unroll
        for (int k = 0; k < MAXSEG; k++) if (k < nseg) { myDesc[k * 32 + lane] = make_uint4(seg[k].qrel, seg[k].rrel, seg[k].len, cfk); cfk += nc[k]; }
    }
    const uint8_t *qb = W.qual + W.qbase;
    for (uint32_t lo = 0; lo < total; lo += CHUNK_CAP) {
        const uint32_t hi = min(total, lo + (uint32_t)CHUNK_CAP);
        uint32_t cfk = cf;
#pragma unroll
        for (int k = 0; k < MAXSEG; k++) if (k < nseg) {
            const uint32_t a = max(cfk, lo), b = min(cfk + nc[k], hi);
            for (uint32_t c = a; c < b; c++) myMap[c - lo] = (uint8_t)(k * 32 + lane);
            cfk += nc[k];
        }
        __syncwarp();
        for (uint32_t f = lo + lane; f < hi; f += 64) {
            const uint32_t f2 = f + 32; const bool has2 = f2 < hi;
            const uint4 d1 = myDesc[myMap[f - lo]];
            const uint4 v1 = ldg_stream(reinterpret_cast<const uint4 *>(qb + (d1.x & ~15u)) + (f - d1.w));
            uint4 d2 = d1, v2 = v1;
            if (has2) { d2 = myDesc[myMap[f2 - lo]]; v2 = ldg_stream(reinterpret_cast<const uint4 *>(qb + (d2.x & ~15u)) + (f2 - d2.w)); }
            process_chunk(W, d1, f, v1, sMaskLo, sMaskHi, t_low, t_hi, acc_sum, acc_cnt);
            if (has2) process_chunk(W, d2, f2, v2, sMaskLo, sMaskHi, t_low, t_hi, acc_sum, acc_cnt);
        }
        __syncwarp();
    }
}

constexpr size_t SMEM_BYTES = (size_t)(2 * A_WORDS + LQ_WORDS) * 4 + 2 * 17 * 16 + NFIRST * 4
                            + (size_t)NWARPS * MAXSEG * 32 * 16 + (size_t)NWARPS * CHUNK_CAP + 64 * 4 + NT + N_STATS * 8 + 16;

__global__ void __launch_bounds__(NT, 4) k_pileup_classify(const KParams P) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t *sA = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *sB = sA + A_WORDS;
    uint32_t *sLQ = sB + A_WORDS;
    uint4 *sMaskLo = reinterpret_cast<uint4 *>(sLQ + LQ_WORDS);
    uint4 *sMaskHi = sMaskLo + 17;
    uint32_t *sFirst = reinterpret_cast<uint32_t *>(sMaskHi + 17);
    uint4 *sDesc = reinterpret_cast<uint4 *>(sFirst + NFIRST);
    uint8_t *sMap = reinterpret_cast<uint8_t *>(sDesc + NWARPS * MAXSEG * 32);
    uint32_t *sScan = reinterpret_cast<uint32_t *>(sMap + NWARPS * CHUNK_CAP);
    uint8_t *sLast = reinterpret_cast<uint8_t *>(sScan + 64);
    unsigned long long *sStats = reinterpret_cast<unsigned long long *>(sLast + NT + ((16 - (NT & 15)) & 15));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t w = P.win_first + blockIdx.x;

    Win W;
    W.wb = (long long)P.region_start + (long long)w * WREAL - 1;
    W.wend = min(W.wb + WN, (long long)P.region_end);
    W.qual = P.qual; W.sA = sA; W.sB = sB; W.sLQ = sLQ;
    W.min_bq = P.min_bq; W.min_mapq = P.min_mapq; W.max_low_mapq = P.max_low_mapq;
    const uint32_t n_ent = (uint32_t)(W.wend - W.wb);      // entries in use, >= 2
    const uint32_t r_lo = P.win_rlo[w], r_hi = P.win_rhi[w];
    W.qbase = 0;
    if (r_hi > r_lo) {
        W.qbase = P.qual_off[r_lo] & ~15ull;
        if (P.qual_off[r_hi] - W.qbase > 0xfffffff0ull || r_hi - r_lo > 65535u) {
            // narrow (16-bit) counters and 32-bit quality offsets cannot represent this window
            if (tid == 0) atomicOr(P.err, (r_hi - r_lo > 65535u) ? ERR_DEPTH : ERR_QUAL_SPAN);
            if (tid == 0) P.win_tab[w] = make_uint2(0, 0);
            return;
        }
    }

    for (int i = tid; i < 2 * A_WORDS + LQ_WORDS; i += NT) sA[i] = 0;
    if (tid < 17) {
        uint32_t lo[4], hi[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t ml = 0, mh = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (4 * q + j < tid) ml |= 0xffu << (8 * j);        // bytes below lo = tid
                if (4 * q + j >= tid) mh |= 0xffu << (8 * j);       // bytes at/above hi = tid
            }
            lo[q] = ml; hi[q] = mh;
        }
        sMaskLo[tid] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        sMaskHi[tid] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    }
    if (tid < NFIRST) sFirst[tid] = P.first_tab[tid];
    if (tid < N_STATS) sStats[tid] = 0;
    __syncthreads();

    // ------------------------------------------------------------------ phase A/B: reads -> counters
    const uint32_t t_low = (P.min_bq & 0x7fu) * 0x01010101u, t_hi = (P.min_bq & 0x80u) ? 0xffffffffu : 0u;
    uint4 *myDesc = sDesc + warp * (MAXSEG * 32);
    uint8_t *myMap = sMap + warp * CHUNK_CAP;
    uint32_t acc_sum = 0, acc_cnt = 0;
    unsigned long long acc_mapq = 0;

    for (uint32_t b0 = r_lo + warp * 32; b0 < r_hi; b0 += NWARPS * 32) {
        const uint32_t r = b0 + lane;
        Seg seg[MAXSEG]; int nseg = 0;
        bool cplx = false;
        int p = 0; uint32_t mq = 0, c0 = 0, c1 = 0, lq = 0; uint64_t q0 = 0;
        if (r < r_hi) {
            const uint32_t fl = P.flag[r];
            c0 = P.cigar_off[r]; c1 = P.cigar_off[r + 1];
            bool live = !(fl & 4u) && c1 > c0;
            if (live && P.read_end) live = (long long)P.read_end[r] > W.wb;
            if (live) {
                p = P.pos[r]; mq = P.mapq[r];
                q0 = P.qual_off[r];
                const uint64_t ql = P.qual_off[r + 1] - q0;
                lq = ql > 0xffffffffull ? 0xffffffffu : (uint32_t)ql;
                if (c1 - c0 > (uint32_t)FAST_OPS) cplx = true;
                else {
                    uint32_t ops[FAST_OPS]; int nm = 0;
#pragma unroll
                    for (int k = 0; k < FAST_OPS; k++) {
                        ops[k] = (c0 + k < c1) ? P.cigar[c0 + k] : 0xfu;        // op 15, len 0: no effect
                        nm += op_is_m(ops[k] & 15u) ? 1 : 0;
                    }
                    if (nm > MAXSEG) cplx = true;
                    else {
                        long long rp = p; uint32_t qp = 0;
                        const bool pass = mq >= W.min_mapq;
#pragma unroll
                        for (int k = 0; k < FAST_OPS; k++) {
                            const uint32_t op = ops[k] & 15u, len = ops[k] >> 4;
                            if (op_is_m(op)) {
                                if (pass && nseg < MAXSEG) { Seg s; if (emit_m(W, rp, qp, len, q0, lq, s)) { seg[nseg < MAXSEG ? nseg : 0] = s; nseg++; } }
                                rp += len; qp += len;
                            } else if (op == 2 || op == 3) rp += len;
                            else if (op == 1 || op == 4) qp += len;
                        }
                        emit_read(W, p, rp, mq, acc_mapq);
                    }
                }
            }
        }
        process_segments(W, nseg, seg, myDesc, myMap, sMaskLo, sMaskHi, t_low, t_hi, acc_sum, acc_cnt, lane);

        // long CIGARs: the whole warp expands one read at a time with prefix sums over 32 ops
        uint32_t cmask = __ballot_sync(FULL, cplx);
        while (cmask) {
            const int src = __ffs(cmask) - 1; cmask &= cmask - 1;
            const long long cp = __shfl_sync(FULL, p, src);
            const uint32_t cmq = __shfl_sync(FULL, mq, src), cc0 = __shfl_sync(FULL, c0, src), cc1 = __shfl_sync(FULL, c1, src);
            const uint32_t clq = __shfl_sync(FULL, lq, src);
            const uint64_t cq0 = __shfl_sync(FULL, (unsigned long long)q0, src);
            const bool cpass = cmq >= W.min_mapq;
            long long rp_carry = cp; uint32_t qp_carry = 0;
            for (uint32_t ob = cc0; ob < cc1; ob += 32) {
                const uint32_t v = (ob + lane < cc1) ? P.cigar[ob + lane] : 0xfu;
                const uint32_t op = v & 15u, len = v >> 4;
                const uint32_t rl = op_ref(op) ? len : 0u, ql = op_qry(op) ? len : 0u;
                uint32_t rs = rl, qs = ql;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    const uint32_t t1 = __shfl_up_sync(FULL, rs, dd), t2 = __shfl_up_sync(FULL, qs, dd);
                    if (lane >= dd) { rs += t1; qs += t2; }
                }
                Seg sg[MAXSEG]; int ns = 0;
                if (cpass && op_is_m(op)) { Seg s; if (emit_m(W, rp_carry + (long long)(rs - rl), qp_carry + (qs - ql), len, cq0, clq, s)) { sg[0] = s; ns = 1; } }
                process_segments(W, ns, sg, myDesc, myMap, sMaskLo, sMaskHi, t_low, t_hi, acc_sum, acc_cnt, lane);
                rp_carry += __shfl_sync(FULL, rs, 31); qp_carry += __shfl_sync(FULL, qs, 31);
                if (rp_carry >= W.wend) break;                           // rest of the read lies right of the window
            }
            if (lane == 0) emit_read(W, cp, rp_carry, cmq, acc_mapq);
        }
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase C: scan, classify, segment
    const uint32_t ebase = tid * PPT;
    uint32_t a[PPT], b[PPT];
#pragma unroll
    for (int k = 0; k < PPT; k++) { a[k] = sA[pidx(ebase + k)]; b[k] = sB[pidx(ebase + k)]; }
#pragma unroll
    for (int k = 1; k < PPT; k++) { a[k] += a[k - 1]; b[k] += b[k - 1]; }
    {
        const uint32_t ta = a[PPT - 1], tb = b[PPT - 1];
        uint32_t ia = ta, ib = tb;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const uint32_t t1 = __shfl_up_sync(FULL, ia, dd), t2 = __shfl_up_sync(FULL, ib, dd);
            if (lane >= dd) { ia += t1; ib += t2; }
        }
        if (lane == 31) { sScan[warp] = ia; sScan[NWARPS + warp] = ib; }
        __syncthreads();
        uint32_t oa = ia - ta, ob = ib - tb;
        for (int j = 0; j < warp; j++) { oa += sScan[j]; ob += sScan[NWARPS + j]; }
#pragma unroll
        for (int k = 0; k < PPT; k++) { a[k] += oa; b[k] += ob; }
    }
    // REF_N bits of this thread's 16 entries
    uint32_t nbits;
    {
        const long long p0 = W.wb + (long long)ebase;
        if (p0 >= 0) { const uint32_t wi = (uint32_t)(p0 >> 5); nbits = __funnelshift_r(P.nmask[wi], P.nmask[wi + 1], (uint32_t)(p0 & 31)); }
        else nbits = P.nmask[0] << 1;
    }
    const uint16_t *sLQ16 = reinterpret_cast<const uint16_t *>(sLQ);
    uint32_t st[PPT];
    uint32_t cnt_pack = 0, covered = 0, sraw = 0, sqc = 0;
    const uint32_t k_first = ebase == 0 ? 1u : 0u;                       // entry 0 is the halo
    const uint32_t k_end = n_ent > ebase ? min((uint32_t)PPT, n_ent - ebase) : 0u;
#pragma unroll
    for (int k = 0; k < PPT; k++) {
        const uint32_t e = ebase + k;
        const uint32_t raw = a[k] & 0xffffu, low = a[k] >> 16;
        const uint32_t qc = b[k] - sLQ16[pidx16(e)];
        const uint32_t fst = raw < (uint32_t)NFIRST ? sFirst[raw] : P.first_tab[raw];
        const bool is_low = raw >= P.min_depth_for_low_mapq && low >= fst;
        uint32_t s;
        if ((nbits >> k) & 1u) s = ST_REF_N;
        else if (raw == 0) s = ST_NO_COVERAGE;
        else if (is_low) s = ST_POOR_MAPQ;
        else if (qc < P.min_depth) s = ST_LOW_COVERAGE;
        else if (P.max_depth > 0 && qc > P.max_depth) s = ST_EXCESSIVE;
        else s = ST_CALLABLE;
        st[k] = s;
        const bool valid = (uint32_t)k >= k_first && (uint32_t)k < k_end;
        if (valid) {
            cnt_pack += 1u << (5 * s);
            covered += raw > 0 ? 1u : 0u; sraw += raw; sqc += qc;
            if (P.dbg_raw) {
                const uint32_t o = (uint32_t)(W.wb + e - P.region_start);
                P.dbg_raw[o] = raw; P.dbg_qc[o] = qc; P.dbg_low[o] = low; P.dbg_state[o] = (uint8_t)s;
            }
        }
    }
    sLast[tid] = (uint8_t)st[PPT - 1];
    __syncthreads();
    // run boundaries
    uint32_t bmask = 0, softmask = 0;
    {
        uint32_t prev = tid > 0 ? sLast[tid - 1] : 0xffu;
#pragma unroll
        for (int k = 0; k < PPT; k++) {
            const bool valid = (uint32_t)k >= k_first && (uint32_t)k < k_end;
            const long long pp = W.wb + (long long)(ebase + k);
            const bool forced = pp == 0 || pp == (long long)P.region_start;
            if (valid && (forced || st[k] != prev)) {
                bmask |= 1u << k;
                if (forced && pp != 0 && st[k] == prev) softmask |= 1u << k;
            }
            prev = st[k];
        }
    }
    {
        const uint32_t nb = __popc(bmask);
        uint32_t inb = nb;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) { const uint32_t t = __shfl_up_sync(FULL, inb, dd); if (lane >= dd) inb += t; }
        if (lane == 31) sScan[16 + warp] = inb;
        __syncthreads();
        uint32_t off = inb - nb, total = 0;
        for (int j = 0; j < NWARPS; j++) { const uint32_t t = sScan[16 + j]; if (j < warp) off += t; total += t; }
        if (tid == 0) {
            const uint32_t base = total ? atomicAdd(P.rec_cursor, total) : 0u;
            sScan[32] = base;
            P.win_tab[w] = make_uint2(base, total);
            if (total && (unsigned long long)base + total > P.rec_cap) atomicOr(P.err, ERR_REC_OVERFLOW);
        }
        __syncthreads();
        uint32_t o = sScan[32] + off;
        uint32_t m = bmask;
        while (m) {
            const int k = __ffs(m) - 1; m &= m - 1;
            uint32_t s = 0;
#pragma unroll
            for (int kk = 0; kk < PPT; kk++) if (kk == k) s = st[kk];
            if (o < P.rec_cap)
                P.rec[o] = (unsigned long long)(uint32_t)(W.wb + ebase + k) | ((unsigned long long)s << 32)
                         | ((unsigned long long)((softmask >> k) & 1u) << 40);
            o++;
        }
    }
    // bins: positions of CALLABLE / POOR_MAPPING_QUALITY / REF_N per stride-sized bin
    if (P.n_bins) {
        const bool any = k_end > k_first;
        const uint32_t c_call = (cnt_pack >> (5 * ST_CALLABLE)) & 31u, c_poor = (cnt_pack >> (5 * ST_POOR_MAPQ)) & 31u, c_refn = cnt_pack & 31u;
        const uint32_t we0 = max(1u, (uint32_t)(warp * 32 * PPT)), we1 = min(n_ent, (uint32_t)((warp + 1) * 32 * PPT));
        if (we1 > we0) {                                                   // warp-uniform
            const uint32_t wbin0 = (uint32_t)(W.wb + we0) / P.stride, wbin1 = (uint32_t)(W.wb + we1 - 1) / P.stride;
            if (wbin0 == wbin1) {        // whole warp inside one bin (the common case: stride >> 512)
                const uint32_t s0 = __reduce_add_sync(FULL, c_call), s1 = __reduce_add_sync(FULL, c_poor), s2 = __reduce_add_sync(FULL, c_refn);
                if (lane == 0) {
                    if (s0) atomicAdd(&P.bins[wbin0], (unsigned long long)s0);
                    if (s1) atomicAdd(&P.bins[P.n_bins + wbin0], (unsigned long long)s1);
                    if (s2) atomicAdd(&P.bins[2 * P.n_bins + wbin0], (unsigned long long)s2);
                }
            } else if (any) {
                const uint32_t tb0 = (uint32_t)(W.wb + ebase + k_first) / P.stride, tb1 = (uint32_t)(W.wb + ebase + k_end - 1) / P.stride;
                if (tb0 == tb1) {
                    if (c_call) atomicAdd(&P.bins[tb0], (unsigned long long)c_call);
                    if (c_poor) atomicAdd(&P.bins[P.n_bins + tb0], (unsigned long long)c_poor);
                    if (c_refn) atomicAdd(&P.bins[2 * P.n_bins + tb0], (unsigned long long)c_refn);
                } else {
#pragma unroll
                    for (int k = 0; k < PPT; k++) {
                        if ((uint32_t)k >= k_first && (uint32_t)k < k_end) {
                            const uint32_t bi = (uint32_t)(W.wb + ebase + k) / P.stride;
                            const uint32_t s = st[k];
                            if (s == ST_CALLABLE) atomicAdd(&P.bins[bi], 1ull);
                            else if (s == ST_POOR_MAPQ) atomicAdd(&P.bins[P.n_bins + bi], 1ull);
                            else if (s == ST_REF_N) atomicAdd(&P.bins[2 * P.n_bins + bi], 1ull);
                        }
                    }
                }
            }
        }
    }
    // per-CTA reduction of the additive counters, then one global atomic per counter
    {
        uint32_t v[11];
#pragma unroll
        for (int s = 0; s < 6; s++) v[s] = (cnt_pack >> (5 * s)) & 31u;
        v[6] = covered; v[7] = sraw; v[8] = acc_sum; v[9] = sqc; v[10] = acc_cnt;
#pragma unroll
        for (int i = 0; i < 11; i++) v[i] = __reduce_add_sync(FULL, v[i]);
        unsigned long long mqs = acc_mapq;
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) mqs += __shfl_xor_sync(FULL, mqs, dd);
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < 6; s++) if (v[s]) atomicAdd(&sStats[S_COUNT0 + s], (unsigned long long)v[s]);
            if (v[6]) atomicAdd(&sStats[S_COVERED], (unsigned long long)v[6]);
            if (v[7]) atomicAdd(&sStats[S_SUMCOV], (unsigned long long)v[7]);
            if (v[8]) atomicAdd(&sStats[S_SUMBQ], (unsigned long long)v[8]);
            if (v[9]) atomicAdd(&sStats[S_QBASES], (unsigned long long)v[9]);
            if (v[10]) atomicAdd(&sStats[S_QBASES_B], (unsigned long long)v[10]);
            if (mqs) atomicAdd(&sStats[S_SUMMAPQ], mqs);
        }
        __syncthreads();
        if (tid < N_STATS && sStats[tid]) atomicAdd(&P.stats[tid * STAT_STRIDE], sStats[tid]);
    }
}

// ---------------------------------------------------------------------------------------------
// Small helper kernels
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lower_bound_pos(const int32_t *pos, uint32_t n, long long key) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if ((long long)pos[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

// candidate read range of every window: reads with pos < window end and pos + max_span > halo position
__global__ void k_window_ranges(const int32_t *pos, uint32_t n_reads, uint32_t region_start, uint32_t region_end,
                                const uint32_t *max_span_ptr, uint32_t w_first, uint32_t n_w, uint32_t *win_rlo, uint32_t *win_rhi) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_w) return;
    const uint32_t max_span = *max_span_ptr;
    const uint32_t w = w_first + i;
    const long long wb = (long long)region_start + (long long)w * WREAL - 1;
    const long long wend = min(wb + WN, (long long)region_end);
    win_rlo[w] = lower_bound_pos(pos, n_reads, wb - (long long)max_span + 1);
    win_rhi[w] = lower_bound_pos(pos, n_reads, wend);
}

// pos + reference span of every read, and the maximum span (long-read mode / max_ref_span == 0)
__global__ void k_read_end(const int32_t *pos, const uint32_t *cigar_off, const uint32_t *cigar, uint32_t r0, uint32_t r1,
                           uint32_t *read_end, uint32_t *max_span) {
    const uint32_t r = r0 + blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t span = 0;
    if (r < r1) {
        for (uint32_t c = cigar_off[r]; c < cigar_off[r + 1]; c++) { const uint32_t v = cigar[c]; if (op_ref(v & 15u)) span += v >> 4; }
        if (read_end) read_end[r] = (uint32_t)pos[r] + span;
    }
    span = __reduce_max_sync(FULL, span);
    if (max_span && (threadIdx.x & 31) == 0 && span) atomicMax(max_span, span);
}

// batch append, step 1: validate ordering of the freshly copied (still batch-relative) columns.
// Device entries r0+1 .. r0+n hold the batch's offsets[1..n]; entry r0 is the previous batch's end.
__global__ void k_validate_batch(const int32_t *pos, const uint32_t *cigar_off, const uint64_t *qual_off, uint32_t r0, uint32_t n,
                                 uint32_t *err) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (pos[r0 + i] < 0 || (r0 + i > 0 && pos[r0 + i] < pos[r0 + i - 1])) atomicOr(err, ERR_UNSORTED);
    const uint32_t cprev = i == 0 ? 0u : cigar_off[r0 + i];
    const uint64_t qprev = i == 0 ? 0ull : qual_off[r0 + i];
    if (cigar_off[r0 + i + 1] < cprev || qual_off[r0 + i + 1] < qprev) atomicOr(err, ERR_OFFSETS);
}
// step 2: rebase entries r0+1 .. r0+n onto the contig-wide payload arrays
__global__ void k_rebase_batch(uint32_t *cigar_off, uint64_t *qual_off, uint32_t r0, uint32_t n, uint32_t cigar_base, uint64_t qual_base) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    cigar_off[r0 + i + 1] += cigar_base;
    qual_off[r0 + i + 1] += qual_base;
}

// ASCII reference -> bit-packed N mask; bases past ref_len (but inside the contig) read as 'N' (mod.rs:79-80)
__global__ void k_nmask_from_ascii(const uint8_t *ref, uint64_t ref_len, uint32_t contig_len, uint32_t *nmask, uint32_t n_words) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool isn = false;
    if (p < ref_len) { const uint8_t c = ref[p]; isn = (c == 'N' || c == 'n'); }
    else if (p < contig_len) isn = true;
    const uint32_t word = __ballot_sync(FULL, isn);
    if ((threadIdx.x & 31) == 0 && (p >> 5) < n_words) nmask[p >> 5] = word;
}

// first_tab[raw] = smallest low in [0, raw+1] with (double)low / (double)raw > fraction  (callable_profiler.rs:100-101)
__global__ void k_first_table(uint32_t *first_tab, double fraction) {
    const uint32_t raw = blockIdx.x * blockDim.x + threadIdx.x;
    if (raw >= 65536u) return;
    if (raw == 0) { first_tab[0] = 0xffffffffu; return; }
    uint32_t lo = 0, hi = raw + 1;                         // predicate is monotone in low (IEEE division is monotone)
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ddiv_rn((double)mid, (double)raw) > fraction) hi = mid; else lo = mid + 1;
    }
    first_tab[raw] = lo;
}

// ---------------------------------------------------------------------------------------------
// Interval compaction: per-window record chunks (arbitrary order) -> sorted clb_interval array
// ---------------------------------------------------------------------------------------------
struct IntervalOut { uint32_t start, end; uint8_t state, soft; uint16_t pad; };

// single block: exclusive scan of per-window record counts
__global__ void k_scan_windows(const uint2 *win_tab, uint32_t n_w, uint32_t *win_out, uint32_t *n_total) {
    __shared__ uint32_t sW[32];
    __shared__ uint32_t sCarry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sCarry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_w; base += blockDim.x) {
        const uint32_t i = base + tid;
        const uint32_t v = i < n_w ? win_tab[i].y : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) { const uint32_t t = __shfl_up_sync(FULL, inc, dd); if (lane >= dd) inc += t; }
        if (lane == 31) sW[warp] = inc;
        __syncthreads();
        uint32_t off = sCarry;
        for (int j = 0; j < warp; j++) off += sW[j];
        if (i < n_w) win_out[i] = off + inc - v;
        __syncthreads();
        if (tid == blockDim.x - 1) sCarry = off + inc;
        __syncthreads();
    }
    if (tid == 0) *n_total = sCarry;
}

// one warp per window: move its records to their sorted place (start/state/soft; end filled next)
__global__ void k_gather_intervals(const unsigned long long *rec, const uint2 *win_tab, const uint32_t *win_out, uint32_t n_w,
                                   IntervalOut *out) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_w) return;
    const uint2 t = win_tab[w];
    const uint32_t o = win_out[w];
    for (uint32_t i = lane; i < t.y; i += 32) {
        const unsigned long long r = rec[t.x + i];
        IntervalOut iv; iv.start = (uint32_t)r; iv.end = 0; iv.state = (uint8_t)(r >> 32); iv.soft = (uint8_t)((r >> 40) & 1u); iv.pad = 0;
        out[o + i] = iv;
    }
}

__global__ void k_fill_ends(IntervalOut *out, const uint32_t *n_total, uint32_t region_end) {
    const uint32_t n = *n_total;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i].end = (i + 1 < n) ? out[i + 1].start : region_end;
}

// padded per-counter lines -> compact [N_STATS] prefix of the counter buffer (bins follow it)
__global__ void k_pack_stats(const unsigned long long *stats_padded, unsigned long long *counters) {
    if (threadIdx.x < N_STATS) counters[threadIdx.x] = stats_padded[threadIdx.x * STAT_STRIDE];
}

}  // namespace clb
