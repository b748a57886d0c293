// clb_fast.cuh -- the fast window kernel of the CallableLoci hot path (sm_100a).
//
// Same reference loop as clb_kernels.cuh (mod.rs:17-42,65-120, callable_profiler.rs:89-155, contig_profiler.rs:47-83,
// histogram_plotter.rs:74-102), restructured for the ordinary window: short CIGARs and fewer than 255 reads over any
// position.  Anything else (long reads, deep piles, a shard's first window) is queued for k_pileup_general, the
// round-1 kernel, so results never depend on which kernel took a window.
//
// What is different from the general kernel, and why (ncu of round 1: issue bound at 0.416 warp instructions per cell):
//   * Every warp runs its own pipeline over sub-batches of <= 32 consecutive reads (one per lane).  The qualities of a
//     sub-batch are ONE contiguous byte range: lane 0 moves it into the warp's stage with a single cp.async.bulk (TMA
//     engine, completion on the warp's own mbarrier) and the CIGAR walk of the 32 reads runs while the copy is in
//     flight.  No per-thread address math for the qualities, no LDG wavefronts, no descriptor pool, no CTA barrier
//     until phase C.
//   * Every thread then streams ITS OWN read's M-segment from shared memory with LDS.128: the segment lives in
//     registers, interior chunks need no byte masks, the loop is 6 instructions per 4 bases.
//   * Base-quality PASS counts go to four packed-u8 arrays selected by (entry of chunk byte 0) mod 4, so the four
//     words of a chunk are added with four plain shared atomics and no funnel shifts; qc depth is the sum of the
//     four arrays (phase C realigns them once per 8 entries), which also removes the M-coverage difference array.
//     A byte cannot overflow because the window is only taken when pos[i] - pos[i - 254] >= max_ref_span for every
//     candidate read, i.e. fewer than 255 reads cover any position.
//   * No halo entry: a window always emits a record for its first position, flagged "window soft", and the
//     interval gather drops it when the previous window ended in the same state.
#pragma once
#include "clb_kernels.cuh"

namespace clb {

constexpr int F_CW = 520;                 // words per alignment-class array (2048 / 4 + 8; 520 % 32 == 8 staggers the banks)
constexpr int F_WSTAGE = CLB_F_WSTAGE;    // staged quality bytes per warp and sub-batch of <= 32 reads (k_window_ranges sizes the sub-batches)
constexpr int F_XCAP = 96;                // second-and-later M-segments per window (streamed from global memory at the end; more: general kernel)
constexpr int F_MAXOPS = 64;              // longest CIGAR walked lane-serially
constexpr int F_LOOKBACK = 254;           // depth proof: pos[i] - pos[i - 254] >= max span  =>  depth <= 254 everywhere
constexpr int F_NFIRST = 256;             // low-MAPQ threshold table entries, one byte each (raw depth <= 254)
#ifndef CLB_F_MINB
#define CLB_F_MINB 4
#endif

// shared memory: A | C (4 class arrays) | one stage per warp | X list | byte masks | first table (u8) | control words | one mbarrier per warp.
// The scratch of phase C (scan, last states, warp stats) reuses the stages, which are idle by then.
constexpr int F_OFF_A = 0;
constexpr int F_OFF_C = F_OFF_A + WN * 4;
constexpr int F_OFF_STAGE = F_OFF_C + 4 * F_CW * 4;
constexpr int F_OFF_X = F_OFF_STAGE + NWARPS * F_WSTAGE;
constexpr int F_OFF_MASK = F_OFF_X + F_XCAP * 8;
constexpr int F_OFF_FIRST = F_OFF_MASK + 2 * 17 * 16;
constexpr int F_OFF_CTL = F_OFF_FIRST + F_NFIRST;
constexpr int F_OFF_BAR = F_OFF_CTL + 16 * 4;
constexpr size_t F_SMEM = F_OFF_BAR + NWARPS * 8;
constexpr int F_OFF_SCAN = F_OFF_STAGE;                        // phase C scratch inside the stages
constexpr int F_OFF_LAST = F_OFF_SCAN + 64 * 4;
constexpr int F_OFF_WSTATS = F_OFF_LAST + NT;
static_assert(F_OFF_STAGE % 16 == 0 && F_WSTAGE % 16 == 0, "bulk copies need 16-byte aligned destinations");
static_assert(F_OFF_WSTATS % 8 == 0 && F_OFF_BAR % 8 == 0 && F_OFF_X % 8 == 0 && F_OFF_MASK % 16 == 0 && F_OFF_FIRST % 16 == 0, "alignment");
static_assert(F_OFF_WSTATS + NWARPS * N_STATS * 8 <= F_OFF_X, "phase C scratch fits the stages");
static_assert(CLB_F_MINB * (F_SMEM + 1024) <= 228 * 1024, "shared memory per SM");
enum { FC_BAIL = 0, FC_XCNT = 1 };

// window table word y: record count | first state << 12 | window-soft first record << 16 | last state << 20
__device__ __forceinline__ uint32_t pack_win_y(uint32_t count, uint32_t first_state, uint32_t wsoft, uint32_t last_state) {
    return count | (first_state << 12) | (wsoft << 16) | (last_state << 20);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// global -> shared bulk copy (TMA engine), completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
    uint4 r = make_uint4(0, 0, 0, 0);
    if (!CLB_SMEM_OK(saddr, 16u)) return r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr));
    return r;
}

// 0x80 in every byte of x that is >= T (unsigned).  t_low = (T & 0x7f) * 0x01010101; BQ_HI = (T >= 128).
template <bool BQ_HI>
__device__ __forceinline__ uint32_t bytes_ge(uint32_t x, uint32_t t_low) {
    const uint32_t H = 0x80808080u;
    const uint32_t d = (x | H) - t_low;                  // bit7 = ((x & 0x7f) >= (T & 0x7f)), no cross-byte borrow
    return (BQ_HI ? (x & d) : (x | d)) & H;
}

// One 16-byte chunk whose bytes all belong to the segment: 6 instructions per word (LOP3, IMAD, LOP3, SHF, IDP.4A, ATOMS).
// acc collects 128 x the sum of the passing qualities (the 0x80 flags are the dot-product weights).
template <bool BQ_HI>
__device__ __forceinline__ void count_chunk(const uint4 v, uint32_t dst, uint32_t t_low, uint32_t &acc) {
    const uint32_t l0 = bytes_ge<BQ_HI>(v.x, t_low), l1 = bytes_ge<BQ_HI>(v.y, t_low);
    const uint32_t l2 = bytes_ge<BQ_HI>(v.z, t_low), l3 = bytes_ge<BQ_HI>(v.w, t_low);
    acc = __dp4a(v.x, l0, __dp4a(v.y, l1, __dp4a(v.z, l2, __dp4a(v.w, l3, acc))));
    red_shared(dst, l0 >> 7); red_shared(dst + 4, l1 >> 7); red_shared(dst + 8, l2 >> 7); red_shared(dst + 12, l3 >> 7);
}
template <bool BQ_HI>
__device__ __forceinline__ void chunk_plain(uint32_t src, uint32_t dst, uint32_t t_low, uint32_t &acc) {
    count_chunk<BQ_HI>(lds128(src), dst, t_low, acc);
}
// First / last chunk of a segment: bytes outside [lo, hi) are taken out of the pass flags.
template <bool BQ_HI>
__device__ __forceinline__ void chunk_edge(uint32_t src, uint32_t dst, uint32_t lo, uint32_t hi, const uint4 *sMaskLo, const uint4 *sMaskHi,
                                           uint32_t t_low, uint32_t &acc) {
    const uint4 v = lds128(src);
    const uint4 ml = sMaskLo[lo], mh = sMaskHi[hi];       // 0xFF in bytes outside [lo, hi)
    const uint32_t l0 = bytes_ge<BQ_HI>(v.x, t_low) & ~(ml.x | mh.x), l1 = bytes_ge<BQ_HI>(v.y, t_low) & ~(ml.y | mh.y);
    const uint32_t l2 = bytes_ge<BQ_HI>(v.z, t_low) & ~(ml.z | mh.z), l3 = bytes_ge<BQ_HI>(v.w, t_low) & ~(ml.w | mh.w);
    acc = __dp4a(v.x, l0, __dp4a(v.y, l1, __dp4a(v.z, l2, __dp4a(v.w, l3, acc))));
    red_shared(dst, l0 >> 7); red_shared(dst + 4, l1 >> 7); red_shared(dst + 8, l2 >> 7); red_shared(dst + 12, l3 >> 7);
}

// First chunk of a segment of two or more chunks (bytes below lo are not part of it) / its last chunk (bytes from hi on are not):
// one mask load each instead of the two of chunk_edge.
template <bool BQ_HI>
__device__ __forceinline__ void chunk_first(uint32_t src, uint32_t dst, uint32_t lo, const uint4 *sMaskLo, uint32_t t_low, uint32_t &acc) {
    const uint4 v = lds128(src);
    const uint4 ml = sMaskLo[lo];
    const uint32_t l0 = bytes_ge<BQ_HI>(v.x, t_low) & ~ml.x, l1 = bytes_ge<BQ_HI>(v.y, t_low) & ~ml.y;
    const uint32_t l2 = bytes_ge<BQ_HI>(v.z, t_low) & ~ml.z, l3 = bytes_ge<BQ_HI>(v.w, t_low) & ~ml.w;
    acc = __dp4a(v.x, l0, __dp4a(v.y, l1, __dp4a(v.z, l2, __dp4a(v.w, l3, acc))));
    red_shared(dst, l0 >> 7); red_shared(dst + 4, l1 >> 7); red_shared(dst + 8, l2 >> 7); red_shared(dst + 12, l3 >> 7);
}
template <bool BQ_HI>
__device__ __forceinline__ void chunk_last(uint32_t src, uint32_t dst, uint32_t hi, const uint4 *sMaskHi, uint32_t t_low, uint32_t &acc) {
    chunk_first<BQ_HI>(src, dst, hi, sMaskHi, t_low, acc);
}

// Stream one M-segment: qs = byte offset of its first quality inside the stage (the stage keeps the 16-byte phase of
// the global column), len bases, rrel = window entry of the first base.  Entry e of class a lives at byte e + 16 - a of
// array a, so a chunk whose byte 0 is entry e0 adds its four words to words (e0 >> 2) + 4 .. + 7 of array e0 & 3.
template <bool BQ_HI>
__device__ __forceinline__ void stream_segment(uint32_t stage_s, uint32_t sC_s, uint32_t qs, uint32_t len, uint32_t rrel,
                                               const uint4 *sMaskLo, const uint4 *sMaskHi, uint32_t t_low, uint32_t &acc) {
    const uint32_t head = qs & 15u;
    const uint32_t nc = (head + len + 15u) >> 4;
    uint32_t src = stage_s + (qs & ~15u);
    const int e0 = (int)rrel - (int)head;                                  // >= -15
    uint32_t dst = sC_s + ((uint32_t)e0 & 3u) * (uint32_t)(F_CW * 4) + (uint32_t)(((e0 >> 2) + 4) * 4);
    if (nc == 1u) { chunk_edge<BQ_HI>(src, dst, head, head + len, sMaskLo, sMaskHi, t_low, acc); return; }
    chunk_first<BQ_HI>(src, dst, head, sMaskLo, t_low, acc);
    uint32_t n_int = nc - 2u;                                              // interior chunks: no masks
    src += 16; dst += 16;
#pragma unroll 1
    for (; n_int >= 2u; n_int -= 2u) {
        const uint4 v0 = lds128(src), v1 = lds128(src + 16);               // both loads in flight before the first count
        count_chunk<BQ_HI>(v0, dst, t_low, acc); count_chunk<BQ_HI>(v1, dst + 16, t_low, acc);
        src += 32; dst += 32;
    }
    if (n_int) { chunk_plain<BQ_HI>(src, dst, t_low, acc); src += 16; dst += 16; }
    chunk_last<BQ_HI>(src, dst, head + len - 16u * (nc - 1u), sMaskHi, t_low, acc);
}

// Same from global memory (the few second-and-later M-segments of a window: their qualities are in L2, one 16-byte
// load per chunk; src0 = 16-byte aligned address of the window's first quality byte, qs relative to it).
template <bool BQ_HI>
__device__ __forceinline__ void stream_segment_global(const uint8_t *src0, uint32_t sC_s, uint32_t qs, uint32_t len, uint32_t rrel,
                                                      const uint4 *sMaskLo, const uint4 *sMaskHi, uint32_t t_low, uint32_t &acc,
                                                      uint32_t c_first, uint32_t c_step, const uint8_t *q_lo, const uint8_t *q_hi) {
    const uint32_t head = qs & 15u;
    const uint32_t nc = (head + len + 15u) >> 4;
    const uint4 *src = reinterpret_cast<const uint4 *>(src0 + (qs & ~15u));
    const int e0 = (int)rrel - (int)head;
    const uint32_t dst0 = sC_s + ((uint32_t)e0 & 3u) * (uint32_t)(F_CW * 4) + (uint32_t)(((e0 >> 2) + 4) * 4);
#pragma unroll 1
    for (uint32_t c = c_first; c < nc; c += c_step) {              // the caller spreads the chunks of a segment over c_step threads
        const uint32_t dst = dst0 + 16u * c;
        if (!CLB_GMEM_OK(src + c, q_lo, q_hi)) continue;
        const uint4 v = ldg_stream(src + c);
        const uint32_t lo = c == 0 ? head : 0u, hi = min(16u, head + len - 16u * c);
        const uint4 ml = sMaskLo[lo], mh = sMaskHi[hi];
        const uint32_t l0 = bytes_ge<BQ_HI>(v.x, t_low) & ~(ml.x | mh.x), l1 = bytes_ge<BQ_HI>(v.y, t_low) & ~(ml.y | mh.y);
        const uint32_t l2 = bytes_ge<BQ_HI>(v.z, t_low) & ~(ml.z | mh.z), l3 = bytes_ge<BQ_HI>(v.w, t_low) & ~(ml.w | mh.w);
        acc = __dp4a(v.x, l0, __dp4a(v.y, l1, __dp4a(v.z, l2, __dp4a(v.w, l3, acc))));
        red_shared(dst, l0 >> 7); red_shared(dst + 4, l1 >> 7); red_shared(dst + 8, l2 >> 7); red_shared(dst + 12, l3 >> 7);
    }
}

// What a thread knows about its read of the next sub-batch: loaded one sub-batch ahead so that the DRAM round trips
// overlap the streaming of the current one.
struct FMeta {
    int ps, pback; uint32_t fl, mq, c0, c1, op0, op1, op2, q0, q1;   // q0, q1: relative to the window's first quality byte
};
// the first three CIGAR ops in one round trip (almost every short read has at most three)
__device__ __forceinline__ void fmeta_ops(const KParams &P, FMeta &M) {
    M.op0 = M.c1 > M.c0 ? P.cigar[M.c0] : 0xfu;
    M.op1 = M.c1 > M.c0 + 1u ? P.cigar[M.c0 + 1u] : 0xfu;
    M.op2 = M.c1 > M.c0 + 2u ? P.cigar[M.c0 + 2u] : 0xfu;
}
__device__ __forceinline__ void fmeta_load(const KParams &P, FMeta &M, uint32_t ic, uint64_t wqx) {
    M.fl = P.flag[ic]; M.c0 = P.cigar_off[ic]; M.c1 = P.cigar_off[ic + 1];
    M.ps = P.pos[ic]; M.mq = P.mapq[ic];
    M.q0 = (uint32_t)(P.qual_off[ic] - wqx); M.q1 = (uint32_t)(P.qual_off[ic + 1] - wqx);   // k_window_ranges checked the window's range fits 32 bits
    M.pback = P.pos[ic >= (uint32_t)F_LOOKBACK ? ic - (uint32_t)F_LOOKBACK : 0u];
}

template <bool BQ_HI, bool DBG>
__global__ void __launch_bounds__(NT, CLB_F_MINB) k_pileup_fast(const KParams P) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t w = P.win_first + blockIdx.x;
    // one table record per window, three independent 16-byte loads (no dependent round trip)
    const uint4 wg = P.win_rec[3 * (size_t)w + 2];                         // x: reads per sub-batch (0: general-path window)
    const uint4 wr = P.win_rec[3 * (size_t)w];
    const ulonglong2 wq = *reinterpret_cast<const ulonglong2 *>(P.win_rec + 3 * (size_t)w + 1);
    const uint32_t G = wg.x;
    if (G == 0) return;
    if (*P.err & (ERR_UNSORTED | ERR_OFFSETS)) return;                     // k_validate_batch refused the columns: do not walk them
#ifdef CLB_PHASE_TIMING      // developer builds only (scripts/phase_timing.py)
#define CLB_FSTAMP(i) do { if (P.timing && threadIdx.x == 0) P.timing[(size_t)w * 8 + (i)] = clock64(); } while (0)
#else
#define CLB_FSTAMP(i) do { } while (0)
#endif
    CLB_FSTAMP(0);

    uint32_t *sA = reinterpret_cast<uint32_t *>(smem_raw + F_OFF_A);
    uint32_t *sC = reinterpret_cast<uint32_t *>(smem_raw + F_OFF_C);
    uint2 *sX = reinterpret_cast<uint2 *>(smem_raw + F_OFF_X);
    const uint4 *sMaskLo = reinterpret_cast<const uint4 *>(smem_raw + F_OFF_MASK);
    const uint4 *sMaskHi = sMaskLo + 17;
    const uint8_t *sFirst = smem_raw + F_OFF_FIRST;
    uint32_t *sScan = reinterpret_cast<uint32_t *>(smem_raw + F_OFF_SCAN);
    uint8_t *sLast = smem_raw + F_OFF_LAST;
    unsigned long long *sWStats = reinterpret_cast<unsigned long long *>(smem_raw + F_OFF_WSTATS);
    volatile uint32_t *sCtl = reinterpret_cast<volatile uint32_t *>(smem_raw + F_OFF_CTL);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t sA_s = smem_addr(sA), sC_s = smem_addr(sC);
    const uint32_t stage_s = smem_addr(smem_raw + F_OFF_STAGE) + (uint32_t)warp * F_WSTAGE;   // this warp's stage
    const uint32_t bar = smem_addr(smem_raw + F_OFF_BAR) + 8u * (uint32_t)warp;                // and its mbarrier
    const uint32_t xcnt_s = smem_addr(const_cast<uint32_t *>(&sCtl[FC_XCNT]));

    const uint32_t wb0 = P.region_start + w * (uint32_t)WREAL;                           // position of entry 0 (BAM positions are below 2^31)
    const uint32_t n_ent = min((uint32_t)WREAL, P.region_end - wb0);                     // entries in use, 1 .. WREAL
    const uint32_t r_lo = wr.x, r_hi = wr.y;
    const uint32_t n_sub = wg.y;                                           // = ceil((r_hi - r_lo) / G), divided once by k_window_ranges;                     // sub-batches of G <= 32 reads; warp v takes v, v + 8, ...
    const uint32_t ebase = tid * PPT;

    // loads whose latency the setup hides: this warp's first sub-batch and the REF_N bits of phase C
    FMeta M;
    uint32_t j = (uint32_t)warp;
    if (j < n_sub) fmeta_load(P, M, min(r_lo + j * G + (uint32_t)lane, r_hi - 1u), wq.x);
    uint32_t nm0, nm1;
    {
        const uint32_t p0 = wb0 + ebase;
        nm0 = P.nmask[p0 >> 5]; nm1 = P.nmask[(p0 >> 5) + 1];
    }
    const uint32_t max_span = r_hi > r_lo ? *P.max_span : 0u;
    // the CIGAR ops are the last hop of the window's first dependent chain (table record -> columns -> ops): start them towards L2 now
    if (tid == 0 && wg.w > wg.z) l2_prefetch(P.cigar, 4ull * wg.z, 4ull * wg.w, 64u << 10);
    {
        uint4 *z = reinterpret_cast<uint4 *>(smem_raw);
        for (int i = tid; i < (F_OFF_STAGE / 16); i += NT) z[i] = make_uint4(0, 0, 0, 0);
        const uint4 *msrc = reinterpret_cast<const uint4 *>(P.win_tables);               // byte masks (first 2 x 17 x 16 bytes)
        uint4 *mdst = reinterpret_cast<uint4 *>(smem_raw + F_OFF_MASK);
        if (tid < 34) mdst[tid] = msrc[tid];
        const uint4 *fsrc = reinterpret_cast<const uint4 *>(P.first_tab8);
        uint4 *fdst = reinterpret_cast<uint4 *>(smem_raw + F_OFF_FIRST);
        if (tid >= 64 && tid < 64 + F_NFIRST / 16) fdst[tid - 64] = fsrc[tid - 64];
        if (tid < 4) sCtl[tid] = 0;
        if (lane == 0) mbar_init(bar, 1);
    }
    __syncthreads();
    CLB_FSTAMP(1);
    if (j < n_sub) fmeta_ops(P, M);

    const uint32_t t_low = (P.min_bq & 0x7fu) * 0x01010101u;
    const uint32_t min_mapq = P.min_mapq, max_low_mapq = P.max_low_mapq;
    uint32_t acc128 = 0;                                                   // 128 x sum of passing qualities of the current sub-batch
    // sums of one window fit 32 bits: at most 2047 positions x 254 reads (depth proof) x 255
    uint32_t acc_sum = 0, acc_mapq = 0;
    uint32_t parity = 0;

    // Every warp runs its own pipeline: bulk copy of its sub-batch's qualities -> CIGAR walk while the copy is in
    // flight -> next sub-batch's columns requested -> wait for the copy -> stream.  Only __syncwarp inside.
    for (; j < n_sub; j += NWARPS) {
        const uint32_t rb = r_lo + j * G, n = min(G, r_hi - rb);
        __syncwarp();                                                      // every lane is done with the stage
        const uint32_t q_first = __shfl_sync(FULL, M.q0, 0), q_last = __shfl_sync(FULL, M.q1, (int)n - 1);
        const uint32_t sb = q_first & ~15u;                                // stage byte 0 <-> this byte of the window's qualities (wq.x is 16-byte aligned)
        const uint32_t bytes = (q_last - sb + 15u) & ~15u;                 // q_last <= 0xfffffff0: no wrap
        const bool fits = bytes <= (uint32_t)F_WSTAGE;                     // k_window_ranges sized the sub-batches: always true
        if (!fits) sCtl[FC_BAIL] = 1;
        if (lane == 0 && fits && bytes) { mbar_arrive_expect_tx(bar, bytes); bulk_g2s(stage_s, P.qual + wq.x + sb, bytes, bar); }

        // ---------------------------------------------------------------- phase A: one read per lane
        bool has0 = false; uint32_t s_qs = 0, s_len = 0, s_rr = 0;
        {
            const bool valid = (uint32_t)lane < n;
            const bool live = valid && !(M.fl & 4u) && M.c1 > M.c0;
            uint32_t nops = live ? M.c1 - M.c0 : 0u;
            const uint32_t ic = rb + (uint32_t)lane;
            const bool deep = valid && ic >= (uint32_t)F_LOOKBACK && (long long)M.pback + (long long)max_span > (long long)M.ps;
            if (deep || nops > (uint32_t)F_MAXOPS) { sCtl[FC_BAIL] = 1; nops = 0; }
            const uint32_t lq = M.q1 - M.q0;
            const uint32_t qs_base = M.q0 - sb;                            // offset of the read's first quality inside the stage
            const uint32_t qx_base = M.q0;                                 // ... and relative to the window's first quality byte
            const int rel = M.ps - (int)wb0;
            const uint32_t mq = M.mq, c0 = M.c0;
            const bool pass = mq >= min_mapq;
            const uint32_t kmax = __reduce_max_sync(FULL, nops);
            int rp = rel; uint32_t qp = 0;
            for (uint32_t k = 0; k < kmax; k++) {
                uint32_t v = k == 0 ? M.op0 : k == 1 ? M.op1 : M.op2;
                if (k >= 3u && k < nops) v = P.cigar[c0 + k];
                if (k >= nops) v = 0xfu;                                                   // op 15, len 0: no effect (skipped reads, shorter CIGARs)
                const uint32_t op = v & 15u, len = v >> 4;
                if (((0x181u >> op) & 1u) && pass && qp < lq && rp < (int)n_ent) {          // M, =, X with qualities, not right of the window
                    const uint32_t l = min(len, lq - qp);
                    const int s = max(rp, 0);
                    const int e = min(rp + (int)l, (int)n_ent);                            // rp < 2048, l < 2^28: no overflow
                    if (e > s) {
                        const uint32_t qo = qp + (uint32_t)(s - rp);
                        if (!has0) { has0 = true; s_qs = qs_base + qo; s_len = (uint32_t)(e - s); s_rr = (uint32_t)s; }
                        else {
                            const uint32_t idx = atom_shared_add(xcnt_s, 1u);
                            if (idx < (uint32_t)F_XCAP) sX[idx] = make_uint2(qx_base + qo, (uint32_t)s | ((uint32_t)(e - s) << 11));
                            else sCtl[FC_BAIL] = 1;
                        }
                    }
                }
                if (((0x18du >> op) & 1u) && rp < (int)WN) rp += (int)len;                  // M, D, N, =, X consume the reference
                if ((0x193u >> op) & 1u) qp += len;                                         // M, I, S, =, X consume the query
            }
            if (nops) {
                const int s = max(rel, 0), e = min(rp, (int)n_ent);
                if (e > s) {
                    const uint32_t delta = 1u + (mq <= max_low_mapq ? 0x10000u : 0u);       // raw depth | low-MAPQ depth << 16
                    red_shared(sA_s + 4u * (uint32_t)s, delta);
                    red_shared(sA_s + 4u * (uint32_t)e, 0u - delta);                        // e <= n_ent <= WREAL < WN
                    if (pass) acc_mapq += mq * (uint32_t)(e - s);
                }
            }
        }
        // this warp's next sub-batch: its columns are in flight while the current one is streamed
        if (j == 0) CLB_FSTAMP(2);
        const bool has_next = j + NWARPS < n_sub;
        if (has_next) fmeta_load(P, M, min(r_lo + (j + NWARPS) * G + (uint32_t)lane, r_hi - 1u), wq.x);
        if (fits && bytes) { mbar_wait(bar, parity); parity ^= 1u; }       // the staged qualities have landed
        if (j == 0) CLB_FSTAMP(3);
        if (has_next) fmeta_ops(P, M);
        // ---------------------------------------------------------------- phase B: stream from shared memory
        if (has0 && fits) stream_segment<BQ_HI>(stage_s, sC_s, s_qs, s_len, s_rr, sMaskLo, sMaskHi, t_low, acc128);
        if (j == 0) CLB_FSTAMP(4);
        acc_sum += acc128 >> 7; acc128 = 0;                                // at most 2047 bases x 255 x 128 per sub-batch: no overflow
    }
    __syncthreads();                                                       // every warp's pushes / counters
    if (sCtl[FC_BAIL]) {
        // not an ordinary window after all: hand it to the general kernel (nothing global was written yet)
        if (tid == 0) P.gen_list[atomicAdd(P.gen_count, 1u)] = w;
        return;
    }
    {
        // second-and-later M-segments of the window's reads (indels): few, pooled for the whole CTA, streamed from L2
        const uint32_t nx = sCtl[FC_XCNT];
        if (nx) {
            for (uint32_t x = (uint32_t)tid >> 4; x < nx; x += NT / 16) {  // 16 threads per segment, one chunk each: one L2 round trip
                const uint2 d = sX[x];
                stream_segment_global<BQ_HI>(P.qual + wq.x, sC_s, d.x, d.y >> 11, d.y & 0x7ffu, sMaskLo, sMaskHi, t_low, acc128,
                                             (uint32_t)tid & 15u, 16u, P.qual, P.qual + P.qual_bytes);
            }
            acc_sum += acc128 >> 7;
            __syncthreads();
        }
    }

    CLB_FSTAMP(5);
    // ------------------------------------------------------------------ phase C: scan, classify, segment
    static_assert(PPT == 8, "phase C of the fast kernel is written for 8 entries per thread");
    uint32_t a[PPT], qcv[PPT];
    {
        const uint4 a0 = *reinterpret_cast<const uint4 *>(sA + ebase), a1 = *reinterpret_cast<const uint4 *>(sA + ebase + 4);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
        // qc depth: entries ebase .. ebase + 7 are bytes ebase + 16 - cls .. of class array cls
        const uint32_t W = 2u * (uint32_t)tid + 4u;
        const uint2 c0v = *reinterpret_cast<const uint2 *>(sC + W);
        uint32_t lo = c0v.x, hi = c0v.y;
#pragma unroll
        for (int cls = 1; cls < 4; cls++) {
            const uint32_t *row = sC + cls * F_CW + W;
            const uint32_t wm = row[-1];
            const uint2 wv = *reinterpret_cast<const uint2 *>(row);
            lo += __funnelshift_r(wm, wv.x, 8u * (4u - (uint32_t)cls));
            hi += __funnelshift_r(wv.x, wv.y, 8u * (4u - (uint32_t)cls));
        }
        qcv[0] = lo & 0xffu; qcv[1] = (lo >> 8) & 0xffu; qcv[2] = (lo >> 16) & 0xffu; qcv[3] = lo >> 24;
        qcv[4] = hi & 0xffu; qcv[5] = (hi >> 8) & 0xffu; qcv[6] = (hi >> 16) & 0xffu; qcv[7] = hi >> 24;
    }
#pragma unroll
    for (int k = 1; k < PPT; k++) a[k] += a[k - 1];
    {
        const uint32_t ta = a[PPT - 1];
        uint32_t ia = ta;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) { const uint32_t t1 = __shfl_up_sync(FULL, ia, dd); if (lane >= dd) ia += t1; }
        if (lane == 31) sScan[warp] = ia;
        __syncthreads();
        const uint32_t oa = ia - ta + __reduce_add_sync(FULL, lane < warp ? sScan[lane] : 0u);
#pragma unroll
        for (int k = 0; k < PPT; k++) a[k] += oa;
    }
    const uint32_t nbits = __funnelshift_r(nm0, nm1, (wb0 + ebase) & 31u);
    const uint32_t k_end = n_ent > ebase ? min((uint32_t)PPT, n_ent - ebase) : 0u;
    const uint32_t vmask = (1u << k_end) - 1u;                              // entries this thread reports
    uint32_t stp = 0;                                                       // 4 bits of state per entry
    uint32_t cnt_pack = 0, covered = 0, sraw = 0, sqc = 0;
    const uint32_t min_dflm = P.min_depth_for_low_mapq, min_depth = P.min_depth;
    const uint32_t max_depth = P.max_depth ? P.max_depth : 0xffffffffu;    // max_depth == 0 disables EXCESSIVE_COVERAGE
#pragma unroll
    for (int k = 0; k < PPT; k++) {
        const uint32_t raw = a[k] & 0xffffu, low = a[k] >> 16;
        const uint32_t qc = qcv[k];
        const uint32_t fst = sFirst[raw];                                   // one byte per depth: raw <= 254 (depth proof)
        const bool is_low = raw >= min_dflm && low >= fst;
        uint32_t s = qc > max_depth ? ST_EXCESSIVE : ST_CALLABLE;
        s = qc < min_depth ? ST_LOW_COVERAGE : s;
        s = is_low ? ST_POOR_MAPQ : s;
        s = raw == 0 ? ST_NO_COVERAGE : s;
        s = ((nbits >> k) & 1u) ? ST_REF_N : s;
        stp |= s << (4 * k);
        cnt_pack += 1u << (5 * s);
        covered += raw > 0 ? 1u : 0u;
        sraw += raw; sqc += qc;
    }
    constexpr uint32_t ALL_ENTRIES = (1u << PPT) - 1u;
    if (vmask != ALL_ENTRIES) {
        // entries past the region end carry no depth (reads are clipped there) but were counted as states: take them out
        uint32_t inv = ~vmask & ALL_ENTRIES;
        while (inv) {
            const int k = __ffs(inv) - 1; inv &= inv - 1;
            cnt_pack -= 1u << (5 * ((stp >> (4 * k)) & 15u));
        }
    }
    if (DBG && P.dbg_raw) {
#pragma unroll
        for (int k = 0; k < PPT; k++) if ((vmask >> k) & 1u) {
            const uint32_t o = wb0 + ebase + k - P.region_start;
            P.dbg_raw[o] = a[k] & 0xffffu; P.dbg_qc[o] = qcv[k]; P.dbg_low[o] = a[k] >> 16; P.dbg_state[o] = (uint8_t)((stp >> (4 * k)) & 15u);
        }
    }
    sLast[tid] = (uint8_t)(stp >> (4 * (PPT - 1)));
    if (k_end && ebase + k_end == n_ent) sScan[3 * NWARPS + 1] = (stp >> (4 * (k_end - 1))) & 15u;   // state of the window's last position
    {
        uint32_t v[10];
#pragma unroll
        for (int s = 0; s < 6; s++) v[s] = (cnt_pack >> (5 * s)) & 31u;
        v[6] = covered; v[7] = sraw; v[8] = acc_sum; v[9] = sqc;
#pragma unroll
        for (int i = 0; i < 10; i++) v[i] = __reduce_add_sync(FULL, v[i]);
        const uint32_t mqs = __reduce_add_sync(FULL, acc_mapq), bqs = v[8];
        if (lane == 0) {
            unsigned long long *ws = sWStats + warp * N_STATS;
#pragma unroll
            for (int s = 0; s < 6; s++) ws[S_COUNT0 + s] = v[s];
            ws[S_COVERED] = v[6]; ws[S_SUMCOV] = v[7]; ws[S_SUMBQ] = bqs; ws[S_QBASES] = v[9];
            ws[S_RESERVED] = 0; ws[S_SUMMAPQ] = mqs;
        }
    }
    __syncthreads();
    if (tid < N_STATS) {
        unsigned long long t = 0;
#pragma unroll
        for (int j = 0; j < NWARPS; j++) t += sWStats[j * N_STATS + tid];
        if (t) atomicAdd(&P.stats[tid * STAT_STRIDE], t);
    }
    // run boundaries: entry e starts a run iff its state differs from entry e - 1; entry 0 always does (window soft)
    uint32_t bmask;
    {
        const uint32_t prev_last = tid > 0 ? sLast[tid - 1] : ((stp & 15u) ^ 1u);
        const uint32_t shifted = (stp << 4) | prev_last;
        uint32_t diff = stp ^ shifted;
        diff |= diff >> 1; diff |= diff >> 2;
        bmask = 0;
#pragma unroll
        for (int k = 0; k < PPT; k++) bmask |= ((diff >> (4 * k)) & 1u) << k;
        bmask &= vmask;
    }
    {
        const uint32_t nb = __popc(bmask);
        uint32_t inb = nb;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) { const uint32_t t = __shfl_up_sync(FULL, inb, dd); if (lane >= dd) inb += t; }
        if (lane == 31) sScan[2 * NWARPS + warp] = inb;
        __syncthreads();
        const uint32_t wt = lane < NWARPS ? sScan[2 * NWARPS + lane] : 0u;
        const uint32_t off = inb - nb + __reduce_add_sync(FULL, lane < warp ? wt : 0u), total = __reduce_add_sync(FULL, wt);
        if (tid == 0) {
            const uint32_t base = atomicAdd(P.rec_cursor, total);           // total >= 1
            sScan[3 * NWARPS] = base;
            P.win_tab[w] = make_uint2(base, pack_win_y(total, stp & 15u, w > 0 ? 1u : 0u, sScan[3 * NWARPS + 1]));
            if ((unsigned long long)base + total > P.rec_cap) atomicOr(P.err, ERR_REC_OVERFLOW);
        }
        __syncthreads();
        uint32_t o = sScan[3 * NWARPS] + off;
        uint32_t m = bmask;
        while (m) {
            const int k = __ffs(m) - 1; m &= m - 1;
            const uint32_t s = (stp >> (4 * k)) & 15u;
            const bool wsoft = w > 0 && tid == 0 && k == 0;
            if (o < P.rec_cap)
                P.rec[o] = (unsigned long long)(wb0 + ebase + k) | ((unsigned long long)s << 32)
                         | ((unsigned long long)(wsoft ? 1u : 0u) << 41);
            o++;
        }
    }
    // bins: positions of CALLABLE / POOR_MAPPING_QUALITY / REF_N per stride-sized bin
    if (P.n_bins) {
        const uint32_t c_call = (cnt_pack >> (5 * ST_CALLABLE)) & 31u, c_poor = (cnt_pack >> (5 * ST_POOR_MAPQ)) & 31u, c_refn = cnt_pack & 31u;
        const uint32_t we0 = (uint32_t)(warp * 32 * PPT), we1 = min(n_ent, (uint32_t)((warp + 1) * 32 * PPT));
        if (we1 > we0) {                                                   // warp-uniform
            const uint32_t fb = wr.z, nbe = wr.w - 1u;                     // first bin; entry where the next bin starts (k_window_ranges counts from the halo)
            uint32_t wbin0, wbin1;
            if (we1 <= nbe) { wbin0 = wbin1 = fb; }
            else if (we0 >= nbe && we1 - nbe <= P.stride) { wbin0 = wbin1 = fb + 1; }
            else { wbin0 = (wb0 + we0) / P.stride; wbin1 = (wb0 + we1 - 1) / P.stride; }
            if (wbin0 == wbin1) {
                const uint32_t s0 = __reduce_add_sync(FULL, c_call), s1 = __reduce_add_sync(FULL, c_poor), s2 = __reduce_add_sync(FULL, c_refn);
                if (lane == 0) {
                    if (s0) atomicAdd(&P.bins[wbin0], (unsigned long long)s0);
                    if (s1) atomicAdd(&P.bins[P.n_bins + wbin0], (unsigned long long)s1);
                    if (s2) atomicAdd(&P.bins[2 * P.n_bins + wbin0], (unsigned long long)s2);
                }
            } else if (k_end) {
                const uint32_t tb0 = (wb0 + ebase) / P.stride, tb1 = (wb0 + ebase + k_end - 1) / P.stride;
                if (tb0 == tb1) {
                    if (c_call) atomicAdd(&P.bins[tb0], (unsigned long long)c_call);
                    if (c_poor) atomicAdd(&P.bins[P.n_bins + tb0], (unsigned long long)c_poor);
                    if (c_refn) atomicAdd(&P.bins[2 * P.n_bins + tb0], (unsigned long long)c_refn);
                } else {
#pragma unroll 1
                    for (uint32_t k = 0; k < k_end; k++) {
                        const uint32_t bi = (wb0 + ebase + k) / P.stride;
                        const uint32_t s = (stp >> (4 * k)) & 15u;
                        if (s == ST_CALLABLE) atomicAdd(&P.bins[bi], 1ull);
                        else if (s == ST_POOR_MAPQ) atomicAdd(&P.bins[P.n_bins + bi], 1ull);
                        else if (s == ST_REF_N) atomicAdd(&P.bins[2 * P.n_bins + bi], 1ull);
                    }
                }
            }
        }
    }
    CLB_FSTAMP(6);
#ifdef CLB_PHASE_TIMING
    if (P.timing && tid == 0) P.timing[(size_t)w * 8 + 7] = (long long)(r_hi - r_lo);
#endif
#undef CLB_FSTAMP
}

}  // namespace clb
