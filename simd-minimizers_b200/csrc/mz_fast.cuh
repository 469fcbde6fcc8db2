// mz_fast.cuh -- W-specialised, register-resident minimizer / syncmer kernel for sm_100a.
//
// Persistent warps pull tiles (32 threads x S windows) from a ticket counter; every warp is an
// autonomous worker (own look-back, own queues), so there is no block barrier after start-up.
// One thread walks S consecutive windows.  Per van-Herk block of W k-mers (fully unrolled):
//   * the words holding the entering- and leaving-base streams were prefetched one block ahead;
//     they are re-aligned with funnel shifts and interleaved so that every byte holds
//     (in,in,out,out) of two consecutive bases;
//   * one LDS.128 from a 256-entry table (requested two pairs of k-mers ahead) returns the
//     rolling-hash deltas of BOTH bases for the forward and the reverse-complement hash
//     (ntHash/mulHash are GF(2)-linear in the tables), so a k-mer hash costs SHF+LOP3 per strand;
//   * (hash & 0xffff0000) | pos goes through a prefix-min / suffix-min pair whose W-entry suffix
//     array lives in registers (static indexing); the rightmost minimum uses max on the
//     complemented key, exactly the reference's packing (src/sliding_min.rs:190-195,336-341);
//   * leftmost != rightmost (1.2e-4 of windows) is tested per group of four windows; the strand
//     rule (src/canonical.rs) runs out of line for those groups only;
//   * a window whose selection differs from the previous window's (src/collect.rs:39-76) pushes ONE
//     16-bit entry -- (selected k-mer << 5 | distance to the window end), built by a single IMAD --
//     onto the lane's queue in shared memory with a predicated store: nothing is recorded for the
//     nine windows in ten that emit nothing, and nothing is re-read or re-walked afterwards.
// The warp then turns its 32 queues into ordered, coalesced output: warp scan of the lane counts,
// decoupled look-back over tile descriptors (resolved one tile later, so predecessors have
// published), and one software-pipelined pass in which lane x & 31 handles output entry x (owner
// lane found by walking the 33 lane offsets, k-mer words fetched one entry ahead).
// Shared memory is planned against the carve-out steps of the SM (194 KB keeps 32 KB of L1 for the
// input stream; dense outputs take all 226 KB), queue rows hold 33 entries so that the rows of one
// lane walk all banks.
#pragma once
#include "../../include/mz_b200.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include "mz_emit.cuh"

namespace mz {

constexpr uint32_t FAST_MAX_W = 32;
// One block of 16 autonomous warps per SM.  (Warps never synchronise after start-up, so the block
// size only decides how many warps share one copy of the hash table in shared memory.)
#ifndef MZ_FAST_NT
#define MZ_FAST_NT 512
#endif
constexpr uint32_t FAST_NT = MZ_FAST_NT;   // threads per block
constexpr uint32_t FAST_BPS = 1;    // resident blocks per SM
// The 256-entry delta table is stored FAST_TC times, copy c = lane & 7 interleaved so that entry
// e of copy c sits at 16-byte slot e*8 + c: the eight lanes of a quarter warp (one LDS.128 pass)
// always hit eight different bank groups -> no bank conflicts on the random table lookups.
#ifndef MZ_FAST_TD
#define MZ_FAST_TD 2
#endif
#ifndef MZ_FAST_TC
#define MZ_FAST_TC 8
#endif
constexpr uint32_t FAST_TC = MZ_FAST_TC;
// Shared memory per block.  The SM splits 228 KB between shared memory and L1 in steps (0, 8, 16, 32,
// 64, 100, 132, 164, 196, 228 KB of shared memory), and the carve-out must hold the dynamic request
// plus 1 KB per block: a request of 205 KB takes the last step and leaves the kernel WITHOUT L1 for
// its input stream (every lane walks its own cache line: 16 warps x 32 lanes x 128 B = 64 KB of
// active lines per SM).  Segments that fit 194 KB stay below the 196 KB step and keep 32 KB of L1
// (C2: 658 -> 714 Gbp/s for the same segments).  Dense outputs, whose segments are cut short by
// the queue space, do better with all of it (closed syncmers w = 11: 482 with 194 KB, 496 with 204
// KB at the same slack): they get up to 226 KB.
constexpr size_t FAST_SMEM_L1 = 194 * 1024;   // leaves 32 KB of L1
constexpr size_t FAST_SMEM_MAX = 226 * 1024;  // no L1 (sm_100: 227 KB per block at most)
inline size_t fast_smem_limit(bool keep_l1) {
    static const size_t env = getenv("MZ_FAST_SMEM_KB") ? (size_t)atoi(getenv("MZ_FAST_SMEM_KB")) * 1024 : 0;
    return env ? env : (keep_l1 ? FAST_SMEM_L1 : FAST_SMEM_MAX);
}
constexpr uint32_t FAST_WARPS = FAST_NT / 32;
// Entries per queue row: the 32 lanes plus one unused entry.  With rows of exactly 64 / 128 bytes the
// rows of one lane sit in only two banks, and the emission pass, where the 32 lanes of a warp read
// 32 consecutive rows of the same owner lane, ran into a 16-way bank conflict on every queue load
// (37 M of the 51 M shared-memory conflicts of an 800 Mbp launch).  Rows of 33 entries walk all
// 32 banks.  (Costs 3 % of the queue space: C2 715 -> 717, forward k=21 w=11 997 -> 1056 Gbp/s.)
constexpr uint32_t FAST_ROWENT = 33;
constexpr uint32_t FAST_TOFFS = 36;  // words per warp and buffer: 33 lane offsets (+ padding)

// Small windows: several van-Herk blocks (B of them, B*W <= 32 k-mers) share one loop iteration, so
// the per-iteration ingest overhead is amortised over ~32 k-mers for every W.
__host__ __device__ constexpr uint32_t fast_b(uint32_t W) { return W <= 16 ? 32 / W : 1; }
__host__ __device__ constexpr uint32_t fast_sb(uint32_t W) { return fast_b(W) * W; }
// Windows longer than FAST_MAX_W (XW instances): the minimum over w k-mers is the minimum over
// T+1 shifted sub-windows of wt = fast_wt(w) <= 24 k-mers each; the kernel instance is the one of
// wt.  For w <= FAST_MAX_W, wt = w.
constexpr uint32_t FAST_XW_MAX_W = 255;  // selected position - window start must fit a byte
__host__ __device__ inline uint32_t fast_wt(uint32_t w) {
    if (w <= FAST_MAX_W) return w;
    const uint32_t parts = (w + 23) / 24;
    return (w + parts - 1) / parts;
}
// Elements (k-mers) a thread computes in front of its first valid window end: the w k-mers of the
// window in front of its segment (whose result only seeds the dedup comparison -- the reference's
// lane-seam rule, src/collect.rs:252-272), rounded up to whole van-Herk blocks so that the
// lead-in ends on a block boundary and everything it pushed is dropped by resetting the queue.
__host__ __device__ inline uint32_t fast_lead(uint32_t w) {
    const uint32_t wt = fast_wt(w);
    return (w + wt - 1) / wt * wt;
}
__host__ __device__ inline uint32_t fast_nb(uint32_t S, uint32_t w) {
    const uint32_t sb = fast_sb(fast_wt(w));
    return (fast_lead(w) + S + sb - 1) / sb;
}
// queue entry: u16 = selected k-mer (11 bits, lane-local) << 5 | window end - selected k-mer;
// long windows need 8 bits for the distance and use u32 entries
__host__ __device__ inline uint32_t fast_qbytes(uint32_t w) { return w > FAST_MAX_W ? 4u : 2u; }
// XW instances: words per warp for the ring of level-1 results (one row of 32 lanes per k-mer;
// leftmost and rightmost sub-window minimum for strand-aware builders)
__host__ __device__ inline uint32_t fast_ring_rows(uint32_t w) {  // power of two >= (w - wt) + iteration
    const uint32_t need = (w - fast_wt(w)) + fast_sb(fast_wt(w)) + 4;
    uint32_t r = 32;
    while (r < need) r <<= 1;
    return r;
}
inline size_t fast_r1_words(uint32_t w, bool lr) {
    return w <= FAST_MAX_W ? 0 : (size_t)fast_ring_rows(w) * 32 * (lr ? 2 : 1);
}
// Global spill area, words per WARP and buffer: every entry a lane can push in one tile.  Only
// touched when a lane's queue overflows (low-complexity sequence: far more minimizers than the
// random-sequence density the queues are sized for).
inline size_t fast_spill_words(uint32_t S, uint32_t w) {
    return (size_t)fast_nb(S, w) * fast_sb(fast_wt(w)) * 32 * fast_qbytes(w) / 4;
}
// shared memory: table | misc | per warp: lane offsets (2 tiles in flight) | queues (2 tiles in flight)
inline size_t fast_smem(uint32_t w, uint32_t q_rows, uint32_t q_bufs = 2) {
    return 256 * FAST_TC * 16 + 64 +
           (size_t)FAST_WARPS * (2 * FAST_TOFFS * 4 + (size_t)q_bufs * q_rows * FAST_ROWENT * fast_qbytes(w));
}
inline double fast_density(const mz_params& p) {
    return p.mode == MZ_MODE_MINIMIZER ? 2.0 / (p.w + 1.0)
         : p.mode == MZ_MODE_CLOSED_SYNCMER ? (p.w == 1 ? 1.0 : 2.0 / p.w) : 1.0 / p.w;
}
// Queue buffers per warp.  Two: the look-back + emission of a tile is deferred until the warp's
// next tile has been computed (its predecessors have published by then).  One, for dense outputs
// (more than about one window in six emits: measured cross-over, w <= 10): the queues then hold twice the segment length, which
// halves the (k + w - 2)-base warm-up per window; the emission pass is long enough there that
// waiting for the predecessors costs less than that.
inline uint32_t fast_q_bufs(const mz_params& p) {
    static const int force = getenv("MZ_FAST_NBUF") ? atoi(getenv("MZ_FAST_NBUF")) : 0;
    if (force == 1 || force == 2) return (uint32_t)force;
    // (syncmers push no lead-in entries that are dropped again and their emission is the longer one
    // with l-mer values: closed syncmers w = 11, density 0.18, measured 422 -> 434 Gbp/s with two buffers)
    return fast_density(p) >= (p.mode == MZ_MODE_MINIMIZER ? 0.175 : 0.19) ? 1u : 2u;
}
// rows a lane may hold before the warp spills: expected entries of a segment + slack.  Queue space is
// what limits the segment length (and with it the warm-up overhead) and decides whether the block
// keeps any L1, so the slack is tight: 3 sigma for minimizers, 2 for syncmers (whose sigma below is
// the Poisson one, an over-estimate).  A lane in ~700 runs over in a tile; the warp then moves its
// queues to the global spill area and that tile is emitted out of line.  Measured (C2 / closed
// syncmers w = 11 + u128 values, Gbp/s): 6 sigma 667 / 411, 4: 710 / 460, 3: 716 / 482, 2: 702 / 499,
// 1.5: 662 / 512, 1: - / 489.
inline uint32_t fast_q_trig(uint32_t S, const mz_params& p) {
    // (the number of minimizers in S windows has a standard deviation of about sqrt(2 S / 3 w):
    // gaps are close to uniform on 1..w; syncmers: Poisson-like)
    static const double nsig_env = getenv("MZ_FAST_QSIGMA") ? atof(getenv("MZ_FAST_QSIGMA")) : -1.0;
    const double nsig = nsig_env >= 0 ? nsig_env : (p.mode == MZ_MODE_MINIMIZER ? 3.0 : 2.0);
    const double mean = S * fast_density(p);
    const double sigma = p.mode == MZ_MODE_MINIMIZER ? sqrt(2.0 * S / (3.0 * p.w)) : sqrt(mean);
    return (uint32_t)std::min<double>(S + 1.0, mean + nsig * sigma + 4.0);
}

// Word `widx` of the packed stream, any index: 0 in front of the buffer; behind it, words of a
// fixed pseudo-random sequence (bases outside the sequence only ever feed windows that are
// discarded, but a repeated word is a 16-periodic sequence whose hashes tie in every window, and
// the lanes of a launch's last tile that run past the end would take the out-of-line strand rule
// for every group of windows: measured 190 us per launch).
__device__ __forceinline__ uint32_t ld_word_any(const uint32_t* seq, uint64_t nwords, int64_t widx) {
    if (widx < 0) return 0u;
    if ((uint64_t)widx >= nwords) {
        uint32_t x = (uint32_t)widx * 0x9E3779B1u;
        x ^= x >> 15;
        x *= 0x85EBCA77u;
        return x ^ (x >> 13);
    }
    return __ldg(seq + widx);
}

// Rare path (leftmost != rightmost minimum): strand rule 2*#TG > l on the window's l bases
// (src/canonical.rs:19-29).  Kept out of line so the unrolled hot loop stays small.
static __device__ __noinline__ bool window_prefers_left(const uint32_t* seq, uint64_t nwords, int64_t bit, uint32_t l) {
    uint32_t bits = 2u * l, cnt = 0;
    while (bits) {
        const int64_t wl = bit >> 5;
        const uint32_t sh = (uint32_t)bit & 31u;
        uint32_t v = __funnelshift_r(ld_word_any(seq, nwords, wl), ld_word_any(seq, nwords, wl + 1), sh) & 0xAAAAAAAAu;
        const uint32_t take = bits < 32u ? bits : 32u;
        if (take < 32u) v &= (1u << take) - 1u;
        cnt += __popc(v);
        bit += take;
        bits -= take;
    }
    return 2u * cnt > l;
}

// Selections of up to four consecutive windows (the last k-mer of the first one is local element
// `pos`) whose leftmost (r) and rightmost (m, complemented key) minimum may differ; bit0 = bit
// position of local base 0.  One call per group of four: the hot loop only carries the call.
static __device__ __noinline__ uint4 tie_fix4(const KArgs& a, int64_t bit0, uint32_t pos, uint32_t ng, uint4 r, uint4 m) {
    uint32_t rr[4] = {r.x, r.y, r.z, r.w};
    const uint32_t mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
    for (uint32_t u = 0; u < 4; u++) {
        if (u >= ng || ((rr[u] ^ mm[u]) & 0xffffu) == 0u) continue;
        const int64_t wbit = bit0 + 2 * ((int64_t)(pos + u) - (int64_t)(a.w - 1u));
        if (!window_prefers_left(a.seq, a.seq_nwords, wbit, a.l)) rr[u] = mm[u] ^ 0xffff0000u;
    }
    return make_uint4(rr[0], rr[1], rr[2], rr[3]);
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
// a*b + c written as mad.lo in PTX (b from a volatile mov when it is 1): the positions and the
// queue entries are built on the FMA pipe, which idles while the ALU pipe is the busy one.
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// Keep a base pointer in registers: without this ptxas re-derives it from kernel parameters in
// front of every access (several 64-bit adds per load / store in the emission loops).
template <typename T>
__device__ __forceinline__ T* pinned(T* p) {
    asm volatile("" : "+l"(p));
    return p;
}
__device__ __forceinline__ uint32_t get_byte(uint32_t w, int J) {
    return J == 0 ? (w & 0xffu) : J == 3 ? (w >> 24) : __byte_perm(w, 0, 0x4440 + J);
}

// ---- per-lane emission queue (shared memory; qa = byte address of the lane's next row) -----------
// push `ent` when a != b: SETP + predicated STS + predicated add, no branch
template <int ROWB, bool WIDE>
__device__ __forceinline__ void q_push_ne(uint32_t& qa, uint32_t a, uint32_t b, uint32_t ent) {
    if constexpr (WIDE)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, %2;\n\t@p st.shared.u32 [%0], %3;\n\t@p add.u32 %0, %0, %4;\n\t}"
                     : "+r"(qa) : "r"(a), "r"(b), "r"(ent), "n"(ROWB));
    else
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b16 lo;\n\tsetp.ne.u32 p, %1, %2;\n\tcvt.u16.u32 lo, %3;\n\t"
                     "@p st.shared.u16 [%0], lo;\n\t@p add.u32 %0, %0, %4;\n\t}"
                     : "+r"(qa) : "r"(a), "r"(b), "r"(ent), "n"(ROWB));
}
// Up to four consecutive windows in one go (NG of them): entry u is pushed when its selection r[u]
// differs from the one before it (prev, r[0], r[1], r[2]).  One asm block, so that the queue
// pointer stays in one register across the four predicated bumps.
// (bump = one * ROWB + qa with ROWB as an immediate: only `one` needs a register)
template <bool WIDE, int NG>
__device__ __forceinline__ void q_push4_ne_fma(uint32_t& qa, uint32_t prev, const uint32_t (&r)[4], const uint32_t (&e)[4],
                                               uint32_t one) {
    if constexpr (WIDE) {
        asm volatile(
            "{\n\t.reg .pred p0, p1, p2, p3;\n\t"
            "setp.ne.u32 p0, %1, %9;\n\tsetp.ne.u32 p1, %2, %1;\n\tsetp.ne.u32 p2, %3, %2;\n\tsetp.ne.u32 p3, %4, %3;\n\t"
            "@p0 st.shared.u32 [%0], %5;\n\t@p0 mad.lo.u32 %0, %10, 132, %0;\n\t"
            "@p1 st.shared.u32 [%0], %6;\n\t@p1 mad.lo.u32 %0, %10, 132, %0;\n\t"
            "@p2 st.shared.u32 [%0], %7;\n\t@p2 mad.lo.u32 %0, %10, 132, %0;\n\t"
            "@p3 st.shared.u32 [%0], %8;\n\t@p3 mad.lo.u32 %0, %10, 132, %0;\n\t}"
            : "+r"(qa)
            : "r"(r[0]), "r"(NG > 1 ? r[1] : r[0]), "r"(NG > 2 ? r[2] : (NG > 1 ? r[1] : r[0])),
              "r"(NG > 3 ? r[3] : (NG > 2 ? r[2] : (NG > 1 ? r[1] : r[0]))), "r"(e[0]), "r"(e[1]), "r"(e[2]), "r"(e[3]),
              "r"(prev), "r"(one));
    } else {
        asm volatile(
            "{\n\t.reg .pred p0, p1, p2, p3;\n\t.reg .b16 l0, l1, l2, l3;\n\t"
            "setp.ne.u32 p0, %1, %9;\n\tsetp.ne.u32 p1, %2, %1;\n\tsetp.ne.u32 p2, %3, %2;\n\tsetp.ne.u32 p3, %4, %3;\n\t"
            "cvt.u16.u32 l0, %5;\n\tcvt.u16.u32 l1, %6;\n\tcvt.u16.u32 l2, %7;\n\tcvt.u16.u32 l3, %8;\n\t"
            "@p0 st.shared.u16 [%0], l0;\n\t@p0 mad.lo.u32 %0, %10, 66, %0;\n\t"
            "@p1 st.shared.u16 [%0], l1;\n\t@p1 mad.lo.u32 %0, %10, 66, %0;\n\t"
            "@p2 st.shared.u16 [%0], l2;\n\t@p2 mad.lo.u32 %0, %10, 66, %0;\n\t"
            "@p3 st.shared.u16 [%0], l3;\n\t@p3 mad.lo.u32 %0, %10, 66, %0;\n\t}"
            : "+r"(qa)
            : "r"(r[0]), "r"(NG > 1 ? r[1] : r[0]), "r"(NG > 2 ? r[2] : (NG > 1 ? r[1] : r[0])),
              "r"(NG > 3 ? r[3] : (NG > 2 ? r[2] : (NG > 1 ? r[1] : r[0]))), "r"(e[0]), "r"(e[1]), "r"(e[2]), "r"(e[3]),
              "r"(prev), "r"(one));
    }
}

// Syncmers, windows <= 32: the low five bits of an entry are the distance d of the selected k-mer
// from the window end, and the window is a syncmer when bit d of `smask` is set (closed: d in
// {0, W-1}; open: d = (W-1)/2; src/syncmers.rs:33-37).  SHF.R.W takes the shift amount mod 32, so
// the test is one shift and one LOP3 with a predicate result per window.  m[u] = smask for the
// group's windows, 0 for the slots behind the last one.
__device__ __forceinline__ void q_push4_mask_fma(uint32_t& qa, const uint32_t (&m)[4], const uint32_t (&e)[4], uint32_t one) {
    asm volatile(
        "{\n\t.reg .pred p0, p1, p2, p3;\n\t.reg .b16 l0, l1, l2, l3;\n\t.reg .b32 t0, t1, t2, t3;\n\t"
        "shf.r.wrap.b32 t0, %1, 0, %5;\n\tshf.r.wrap.b32 t1, %2, 0, %6;\n\tshf.r.wrap.b32 t2, %3, 0, %7;\n\tshf.r.wrap.b32 t3, %4, 0, %8;\n\t"
        "and.b32 t0, t0, 1;\n\tand.b32 t1, t1, 1;\n\tand.b32 t2, t2, 1;\n\tand.b32 t3, t3, 1;\n\t"
        "setp.ne.u32 p0, t0, 0;\n\tsetp.ne.u32 p1, t1, 0;\n\tsetp.ne.u32 p2, t2, 0;\n\tsetp.ne.u32 p3, t3, 0;\n\t"
        "cvt.u16.u32 l0, %5;\n\tcvt.u16.u32 l1, %6;\n\tcvt.u16.u32 l2, %7;\n\tcvt.u16.u32 l3, %8;\n\t"
        "@p0 st.shared.u16 [%0], l0;\n\t@p0 mad.lo.u32 %0, %9, 66, %0;\n\t"
        "@p1 st.shared.u16 [%0], l1;\n\t@p1 mad.lo.u32 %0, %9, 66, %0;\n\t"
        "@p2 st.shared.u16 [%0], l2;\n\t@p2 mad.lo.u32 %0, %9, 66, %0;\n\t"
        "@p3 st.shared.u16 [%0], l3;\n\t@p3 mad.lo.u32 %0, %9, 66, %0;\n\t}"
        : "+r"(qa)
        : "r"(m[0]), "r"(m[1]), "r"(m[2]), "r"(m[3]), "r"(e[0]), "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(one));
}

// entry at shared address `addr`
template <bool WIDE>
__device__ __forceinline__ uint32_t q_load(uint32_t addr) {
    if constexpr (WIDE) {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
        return v;
    } else {
        uint16_t h;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(addr));
        return h;
    }
}
// move n rows of a lane's queue (first row at shared address qa0) to its slice of the spill area
template <int ROWB, bool WIDE, typename QT>
static __device__ __noinline__ void q_spill(uint32_t qa0, uint32_t n, QT* dst) {
    for (uint32_t i = 0; i < n; i++) dst[i] = (QT)q_load<WIDE>(qa0 + i * (uint32_t)ROWB);
}
template <int ROWB, bool WIDE>
__device__ __forceinline__ void q_push_if(uint32_t& qa, bool c, uint32_t ent) {
    q_push_ne<ROWB, WIDE>(qa, (uint32_t)c, 0u, ent);
}

// What one thread of a tile works on.  Local element e is the k-mer starting at base F + e, where
// F = (first window of the segment) - 1 - (lead - w): every thread starts one window (plus block
// padding for long windows) in front of its segment, so the first valid window ends at element
// `lead` for every thread.  F is negative for the first thread of a sequence / a read: bases in
// front of the buffer read as 'A' and only reach discarded windows.
struct FSeg {
    int64_t bit0;        // bit position (in a.seq, may be negative) of local base 0
    uint32_t nvalid;     // windows this thread may emit
    uint32_t first_always;  // no window in front of the segment: its first window always emits
};
__device__ __forceinline__ FSeg make_fseg(const KArgs& a, uint32_t tile, uint32_t lane) {
    FSeg s;
    const int64_t back = 1 + (int64_t)(a.lead - a.w);
    if (a.n_reads == 0) {
        const uint64_t j0 = a.wbeg + ((uint64_t)tile * 32u + lane) * a.S;
        const uint64_t left = j0 < a.wend ? a.wend - j0 : 0;
        s.nvalid = (uint32_t)(left < a.S ? left : a.S);
        s.first_always = (j0 == 0);
        s.bit0 = 2 * ((int64_t)j0 - back) + a.bitbias;
    } else {
        const uint64_t pi = (uint64_t)tile * 32u + lane;
        s.nvalid = 0;
        s.first_always = 1;
        s.bit0 = a.bitbias;
        if (pi < a.n_reads) {
            const uint64_t r = a.piece_read ? a.piece_read[pi] : pi;
            const uint32_t win0 = a.piece_read ? a.piece_win0[pi] : 0u;
            const uint64_t startbits = a.read_start_bp ? 2 * a.read_start_bp[r] : r * a.stride_bits;
            const uint32_t len = a.read_len_bp ? a.read_len_bp[r] : a.fixed_len_bp;
            const uint32_t nw = len >= a.l ? len - a.l + 1 : 0;
            const uint32_t left = nw > win0 ? nw - win0 : 0;
            s.nvalid = left < a.S ? left : a.S;
            s.first_always = (win0 == 0);
            s.bit0 = (int64_t)startbits + 2 * ((int64_t)win0 - back) + a.bitbias;
        }
    }
    return s;
}

// ---- ordered emission of one tile (called by a whole warp, once per tile) -------------------------
// tp[t] = exclusive output offset of lane t inside the tile, tp[32] = the tile's total; entry i of
// lane t lives in the queue row i of the tile (shared memory).  Lane x & 31 handles output entry x;
// the owner lane of x is tracked incrementally (x grows by 32 per step and a lane holds ~40 entries,
// so the owner advances by 0 or 1 almost always).

// (a & c) | (~b & ~c) in one LOP3
__device__ __forceinline__ uint32_t lop3_sel_not(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xB1;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// swap the two bits of every base and complement it (code ^ 2): bit pairs (hi, lo) -> (~lo, hi)
__device__ __forceinline__ uint32_t swap_comp32(uint32_t x) { return lop3_sel_not(x >> 1, x << 1, 0x55555555u); }
// 2-bit reverse complement of the low 2*len bits of (vh:vl) (len <= 32), shifted down; rsh = 64 - 2*len
__device__ __forceinline__ uint64_t revcomp64(uint32_t vl, uint32_t vh, uint32_t rsh) {
    const uint32_t lo = swap_comp32(__brev(vh)), hi = swap_comp32(__brev(vl));  // bit reversal swaps the halves
    return (((uint64_t)hi << 32) | lo) >> rsh;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// One instance per (value width, entry format), so that the loop carries no run-time switches.
// FMT 0: minimizer positions; 1: minimizer positions + super-k-mer starts; 2: syncmers (window start, l-mer).
// The three words of a k-mer (five of an l-mer) are loaded unclamped: tiles so close to the end of
// the buffer that a load could reach behind it take the out-of-line path (fast_emit_generic).
template <int VB, bool XW, int FMT, bool CANON>
__device__ __forceinline__ void fast_emit_tile(const KArgs& a, uint32_t tile, unsigned long long gbase, uint32_t total,
                                                   uint32_t tp_s, uint32_t qs) {
    // qs = shared address of the tile's queue rows: entry i of lane t at qs + i * ROWB + t * sizeof(QT)
    // tp_s = shared address of the 33 lane offsets
    using QT = typename std::conditional<XW, uint32_t, uint16_t>::type;
    constexpr uint32_t DBITS = XW ? 8 : 5, DMASK = (1u << DBITS) - 1u, ROWB = FAST_ROWENT * sizeof(QT);
    constexpr int NW = VB == 64 ? 3 : VB == 128 ? 5 : 0;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t Wr = a.w, S = a.S, len = a.val_len;
    uint32_t* const opos = pinned(a.pos + gbase);
    uint32_t* const osk = pinned(a.sk + (FMT == 1 ? gbase : 0ull));
    unsigned long long* const oval = pinned(reinterpret_cast<unsigned long long*>(a.val) + (VB ? gbase * (VB / 64) : 0ull));
    // Tile-uniform addressing.  Lane t's local element 0 is the k-mer at position pos00 + t * S with
    // pos00 = (first window of the tile) - 1 - (lead - w).
    const uint64_t j00 = a.wbeg + (uint64_t)tile * 32u * S;
    const int64_t back = 1 + (int64_t)(a.lead - Wr);
    const uint32_t pos00 = (uint32_t)((int64_t)j00 - back);
    const int64_t tbit0 = 2 * ((int64_t)j00 - back) + a.bitbias;
    const uint32_t* const twbase = pinned(a.seq + (tbit0 >> 5));  // never dereferenced below word 0
    const uint32_t tsh = (uint32_t)tbit0 & 31u;
    const uint32_t mlo = len < 16 ? (1u << (2 * len)) - 1u : 0xffffffffu;
    const uint32_t mhi = len <= 16 ? 0u : len < 32 ? (1u << (2 * len - 32)) - 1u : 0xffffffffu;
    const uint32_t rsh = 64u - 2u * (len < 32u ? len : 32u);
    const uint32_t wback = Wr - 1u;
    // 128-bit values: word masks of the 2*len valid bits, and the shift 128 - 2*len of the reverse complement
    uint32_t vm[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int nb = (int)(2u * len) - 32 * q;  // valid bits in word q
        vm[q] = nb <= 0 ? 0u : nb >= 32 ? 0xffffffffu : (1u << nb) - 1u;
    }
    const uint32_t rs128 = 128u - 2u * (len < 64u ? len : 64u), rsw = rs128 >> 5, rsb = rs128 & 31u;

    struct Ent {
        uint32_t pos, sk, bp, w[NW ? NW : 1];
    };
    // owner of the entry being located: lane t, tS = t * S, qb = qs + t * sizeof(QT) - tp[t] * ROWB
    uint32_t t = 0, tS = 0, qb = qs, hi = lds32(tp_s + 4u);
    const uint32_t last = total - 1u;
    // locate entry x (< total), decode it, request the words of its k-mer
    auto stageB = [&](uint32_t x, Ent& e) {
        if (x >= hi) {
            uint32_t lo;
            do {
                t++;
                lo = hi;
                hi = lds32(tp_s + 4u * t + 4u);
            } while (x >= hi);
            qb = qs + t * (uint32_t)sizeof(QT) - lo * ROWB;
            tS = t * S;
        }
        const uint32_t v = q_load<XW>(qb + x * ROWB);
        const uint32_t posl = v >> DBITS;  // selected k-mer
        uint32_t rel = tS + posl;          // reported k-mer, counted from the tile's element 0
        if (FMT != 0) {
            const uint32_t wst = rel + (v & DMASK) - wback;  // first k-mer of the window
            e.sk = pos00 + wst;                               // = index of the window
            if (FMT == 2) rel = wst;
        }
        e.pos = pos00 + rel;
        e.bp = tsh + 2u * rel;
        if (NW) {
            const uint32_t* const wp = twbase + (e.bp >> 5);
#pragma unroll
            for (int q = 0; q < NW; q++) e.w[q] = __ldg(wp + q);
        }
    };
    auto stageC = [&](uint32_t x, const Ent& e, auto tail) {
        if (decltype(tail)::value && x >= total) return;  // the last batch of 32 may be partial
        __stcs(opos + x, e.pos);          // streaming stores: the outputs are not read again here
        if (FMT == 1) __stcs(osk + x, e.sk);
        const uint32_t sh = e.bp & 31u;
        if (VB == 64) {
            const uint32_t vl = __funnelshift_r(e.w[0], e.w[NW > 1 ? 1 : 0], sh) & mlo;
            const uint32_t vh = __funnelshift_r(e.w[NW > 1 ? 1 : 0], e.w[NW > 2 ? 2 : 0], sh) & mhi;
            uint64_t v = ((uint64_t)vh << 32) | vl;
            if (CANON) {
                const uint64_t r = revcomp64(vl, vh, rsh);
                v = r < v ? r : v;
            }
            __stcs(oval + x, (unsigned long long)v);
        } else if (VB == 128) {
            // the l-mer as four 32-bit words (low to high), masked to 2*len bits
            uint32_t v[4];
#pragma unroll
            for (int q = 0; q < 4; q++) v[q] = __funnelshift_r(e.w[NW > q ? q : 0], e.w[NW > q + 1 ? q + 1 : 0], sh) & vm[q];
            uint32_t o0 = v[0], o1 = v[1], o2 = v[2], o3 = v[3];
            if (CANON) {
                // reverse the 128 bits two at a time and complement (word q of the result comes from
                // word 3 - q), then shift down by s = 128 - 2*len = 32 * rsw + rsb (tile-uniform)
                const uint32_t r0 = swap_comp32(__brev(v[3])), r1 = swap_comp32(__brev(v[2]));
                const uint32_t r2 = swap_comp32(__brev(v[1])), r3 = swap_comp32(__brev(v[0]));
                uint32_t q0, q1, q2, q3;
                if (rsw == 0) {
                    q0 = __funnelshift_r(r0, r1, rsb), q1 = __funnelshift_r(r1, r2, rsb), q2 = __funnelshift_r(r2, r3, rsb), q3 = r3 >> rsb;
                } else if (rsw == 1) {
                    q0 = __funnelshift_r(r1, r2, rsb), q1 = __funnelshift_r(r2, r3, rsb), q2 = r3 >> rsb, q3 = 0u;
                } else if (rsw == 2) {
                    q0 = __funnelshift_r(r2, r3, rsb), q1 = r3 >> rsb, q2 = 0u, q3 = 0u;
                } else {
                    q0 = r3 >> rsb, q1 = 0u, q2 = 0u, q3 = 0u;
                }
                const uint64_t qh = ((uint64_t)q3 << 32) | q2, vh = ((uint64_t)v[3] << 32) | v[2];
                const uint64_t ql = ((uint64_t)q1 << 32) | q0, vl = ((uint64_t)v[1] << 32) | v[0];
                if (qh < vh || (qh == vh && ql < vl)) o0 = q0, o1 = q1, o2 = q2, o3 = q3;
            }
            // 16-byte streaming store (consecutive lanes -> 512 contiguous bytes per warp)
            asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(oval + 2ull * x), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
        }
    };
    // software pipeline over batches of 32 entries, unrolled twice (two register sets instead of
    // copies): entry x + 32 is located and its words are requested while entry x is turned into a
    // value and stored.  The batch loop is warp-uniform: lanes behind the last entry re-do the last one.
    // (Two entries per lane and step -- x and x + 32 as independent chains, four register sets -- was
    // measured: C2 657 -> 640, C4 437 -> 416 Gbp/s, only w <= 4 gained 2-4 %; not kept.)
    const uint32_t nbatch = (total + 31u) >> 5;
    const std::integral_constant<bool, false> full;
    const std::integral_constant<bool, true> tail;
    Ent e0, e1;
    uint32_t x = lane, i = 1;
    stageB(min(x, last), e0);
#pragma unroll 1
    for (;;) {
        if (i >= nbatch) {
            stageC(x, e0, tail);
            break;
        }
        stageB(min(x + 32u, last), e1);
        stageC(x, e0, full);
        i++;
        if (i >= nbatch) {
            stageC(x + 32u, e1, tail);
            break;
        }
        stageB(min(x + 64u, last), e0);
        stageC(x + 32u, e1, full);
        i++;
        x += 64u;
    }
}

// May a k-mer word load of this tile reach behind the buffer?  (conservative bound on the tile's
// last element: 32 lanes x S windows + lead-in + one loop iteration; 5 words per l-mer)
__device__ __forceinline__ bool fast_tile_at_buffer_end(const KArgs& a, uint32_t tile) {
    const uint64_t j00 = a.wbeg + (uint64_t)tile * 32u * a.S;
    const int64_t tbit0 = 2 * ((int64_t)j00 - 1 - (int64_t)(a.lead - a.w)) + a.bitbias;
    const int64_t wmax = (tbit0 >> 5) + ((31u + 2u * (32u * a.S + a.lead + 64u)) >> 5) + 5;
    return wmax >= (int64_t)a.seq_nwords;
}
// Run-time selection of the emission instance (warp-uniform, once per tile).
template <bool XW, bool SYNC, bool CANON>
__device__ __forceinline__ void fast_emit_seq(const KArgs& a, uint32_t tile, unsigned long long gbase, uint32_t total,
                                              uint32_t tp_s, uint32_t qs) {
    constexpr int F0 = SYNC ? 2 : 0, F1 = SYNC ? 2 : 1;
    if (SYNC || !a.want_sk) {
        if (a.value_bits == 64) fast_emit_tile<64, XW, F0, CANON>(a, tile, gbase, total, tp_s, qs);
        else if (a.value_bits == 128) fast_emit_tile<128, XW, F0, CANON>(a, tile, gbase, total, tp_s, qs);
        else fast_emit_tile<0, XW, F0, false>(a, tile, gbase, total, tp_s, qs);
    } else {
        if (a.value_bits == 64) fast_emit_tile<64, XW, F1, CANON>(a, tile, gbase, total, tp_s, qs);
        else if (a.value_bits == 128) fast_emit_tile<128, XW, F1, CANON>(a, tile, gbase, total, tp_s, qs);
        else fast_emit_tile<0, XW, F1, false>(a, tile, gbase, total, tp_s, qs);
    }
}

// Out-of-line emission for everything that is not the hot single-sequence case: batch mode
// (positions relative to the read / piece that owns the entry) and tiles whose queues were spilled.
// Entry i of lane t lives at ebase + i * sI + t * sT (shared-memory queue rows or the spill area).
template <int VB, bool XW>
static __device__ __noinline__ void fast_emit_generic(const KArgs& a, uint32_t tile, unsigned long long gbase, uint32_t total,
                                                    const uint32_t* tp, const unsigned char* ebase, uint32_t sI, uint32_t sT) {
    using QT = typename std::conditional<XW, uint32_t, uint16_t>::type;
    constexpr uint32_t DBITS = XW ? 8 : 5, DMASK = (1u << DBITS) - 1u;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t Wr = a.w;
    const bool minim = a.mode == MODE_MINIMIZER, want_sk = a.want_sk != 0, canon_val = a.val_canonical != 0;
    uint32_t* const opos = a.pos + gbase;
    uint32_t* const osk = a.sk + (want_sk ? gbase : 0ull);
    unsigned long long* const oval = reinterpret_cast<unsigned long long*>(a.val) + (VB ? gbase * (VB / 64) : 0ull);
    const int64_t back = 1 + (int64_t)(a.lead - Wr);
    uint32_t t = 0, lo = 0, hi = tp[1];
#pragma unroll 1
    for (uint32_t x = lane; x < total; x += 32) {
        while (x >= hi) {
            t++;
            lo = hi;
            hi = tp[t + 1];
        }
        const uint32_t v = *reinterpret_cast<const QT*>(ebase + (size_t)(x - lo) * sI + (size_t)t * sT);
        const uint32_t posl = v >> DBITS, wst = posl + (v & DMASK) - (Wr - 1u);
        uint64_t startbits = 0;
        int64_t first;  // reported position of local element 0 (relative to the read in batch mode)
        if (a.n_reads == 0) {
            first = (int64_t)(a.wbeg + ((uint64_t)tile * 32u + t) * a.S) - back;
        } else {
            const uint64_t pi = (uint64_t)tile * 32u + t;
            const uint64_t r = a.piece_read ? a.piece_read[pi] : pi;
            const uint32_t win0 = a.piece_read ? a.piece_win0[pi] : 0u;
            startbits = a.read_start_bp ? 2 * a.read_start_bp[r] : r * a.stride_bits;
            first = (int64_t)win0 - back;
        }
        const uint32_t p = (uint32_t)(first + (minim ? posl : wst));
        opos[x] = p;
        if (want_sk) osk[x] = (uint32_t)(first + wst);
        const uint64_t bit = (uint64_t)((int64_t)startbits + 2 * (int64_t)p + a.bitbias);
        if (VB == 64) {
            oval[x] = kmer_value_u64(a, bit, a.val_len, canon_val);
        } else if (VB == 128) {
            uint64_t vlo, vhi;
            kmer_value_u128(a, bit, a.val_len, canon_val, vlo, vhi);
            reinterpret_cast<ulonglong2*>(oval)[x] = make_ulonglong2(vlo, vhi);
        }
    }
}

// AMB: windows that contain an ambiguous base (a.amb, one bit per base) produce nothing
// (run_skip_ambiguous_windows, src/lib.rs:451-496); a separate instance so that the plain path
// carries no extra state.
// XW: a.w > FAST_MAX_W.  The van-Herk machinery computes the minima of sub-windows of W k-mers;
// they go through a small per-warp ring of rows in the L2 scratch, and the minimum of the real
// window is the minimum over the T+1 shifted sub-windows that cover it (the current one from
// registers, the others read back from the ring a group of four k-mers ahead of use).  Everything
// downstream (tie test, queue, emission) sees the real-window result.
template <int W, bool HC, bool LR, bool SYNC, bool AMB = false, bool XW = false>
__global__ void __launch_bounds__(FAST_NT, FAST_BPS) mz_fast_kernel(const __grid_constant__ KArgs a) {
    static_assert(W >= 1 && W <= (int)FAST_MAX_W, "W out of range");
    constexpr int B = (int)fast_b(W);    // van-Herk blocks per loop iteration
    constexpr int SB = B * W;            // k-mers per loop iteration (<= 32)
    constexpr uint32_t NT = FAST_NT;
    using QT = typename std::conditional<XW, uint32_t, uint16_t>::type;
    constexpr int ROWB = (int)FAST_ROWENT * (int)sizeof(QT);  // bytes per queue row
    constexpr uint32_t DBITS = XW ? 8 : 5, DMUL = (1u << DBITS) - 1u;  // entry = pos << DBITS | (window end - pos)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint4* T = reinterpret_cast<uint4*>(smem_raw);
    uint32_t* misc = reinterpret_cast<uint32_t*>(T + 256 * FAST_TC);
    const uint32_t QROWS = a.q_rows;
    // this warp's lane offsets and queues, two tiles in flight
    uint32_t* const tp0 = misc + 16 + warp * (2 * FAST_TOFFS);
    const bool defer = a.q_bufs == 2;  // look-back + emission one tile later (two queue buffers)
    unsigned char* const q0 = reinterpret_cast<unsigned char*>(misc + 16 + FAST_WARPS * 2 * FAST_TOFFS) + (size_t)warp * a.q_bufs * QROWS * ROWB;
    const uint32_t q0s = (uint32_t)__cvta_generic_to_shared(q0) + lane * (uint32_t)sizeof(QT);
    const uint32_t Wr = XW ? a.w : (uint32_t)W;  // real window length
    const uint32_t lead = a.lead, NB = a.nb;
    // spill area (global, per warp and buffer) and the XW ring
    const size_t spill_words = a.scratch_words_per_block;
    uint32_t* const sc0 = a.scratch + ((size_t)blockIdx.x * FAST_WARPS + warp) * (2 * spill_words + a.r1_words_per_warp);
    uint32_t* const r1 = sc0 + 2 * spill_words;  // XW: level-1 rows of this warp
    const uint32_t SPILLCAP = NB * (uint32_t)SB;  // entries per lane in the spill area

    const uint32_t k = a.k, R = a.rot & 31u, R2 = (2u * R) & 31u;
    // ---- table: index byte = in0 | in1<<2 | out0<<4 | out1<<6 (two consecutive bases) --------
    // (forward hash: 8-byte entries read with LDS.64, one pass per half warp -> 16 copies)
    constexpr uint32_t TCOPIES = HC ? FAST_TC : 2 * FAST_TC;
    uint2* const T2 = reinterpret_cast<uint2*>(T);
    for (uint32_t slot = tid; slot < 256 * TCOPIES; slot += NT) {
        const uint32_t idx = slot / TCOPIES;  // slot = entry * TCOPIES + copy
        const uint32_t in0 = idx & 3u, in1 = (idx >> 2) & 3u, out0 = (idx >> 4) & 3u, out1 = idx >> 6;
        const uint32_t rk = (R * k) & 31u, rk1 = (R * (k - 1)) & 31u;
        const uint32_t f0 = a.f[in0] ^ rotl32(a.f[out0], rk), f1 = a.f[in1] ^ rotl32(a.f[out1], rk);
        const uint32_t c0 = rotl32(a.c[in0], rk1) ^ rotr32(a.c[out0], R);
        const uint32_t c1 = rotl32(a.c[in1], rk1) ^ rotr32(a.c[out1], R);
        if (HC) T[slot] = make_uint4(f0, rotl32(f0, R) ^ f1, c0, rotr32(c0, R) ^ c1);
        else T2[slot] = make_uint2(f0, rotl32(f0, R) ^ f1);
    }
    if (tid == 0) {
        uint32_t fa = 0, ca = 0;  // hash state of the virtual all-'A' k-mer before every segment
        for (uint32_t j = 0; j < k; j++) {
            fa ^= rotl32(a.f[0], (R * j) & 31u);
            ca ^= rotl32(a.c[0], (R * j) & 31u);
        }
        misc[1] = fa;
        misc[2] = ca;
    }
    const uint32_t tcopy = lane & (TCOPIES - 1u);
    const uint32_t tb = (uint32_t)__cvta_generic_to_shared(T) + tcopy * (HC ? 16u : 8u);  // this lane's copy
    auto table = [&](uint32_t idx) -> uint4 {  // entry idx of this lane's copy (prologue only)
        if (HC) return T[idx * TCOPIES + tcopy];
        const uint2 e = T2[idx * TCOPIES + tcopy];
        return make_uint4(e.x, e.y, 0u, 0u);
    };
    // opaque 1 for imad(): a kernel argument, so that ptxas cannot fold x * 1 + c back into an
    // ALU-pipe add (the positions and the queue pointer are bumped on the idle FMA pipe)
    const uint32_t one = a.one;
    static_assert(ROWB == (XW ? 132 : 66), "q_push4_ne_fma / q_push4_mask_fma hard-code the row size");
    // syncmer offsets d = (window end) - (selected pos): closed {0, W-1}, open {(W-1)/2}
    const uint32_t so1 = a.mode == MODE_CLOSED ? 0u : (Wr - 1) / 2, so2 = a.mode == MODE_CLOSED ? Wr - 1 : (Wr - 1) / 2;
    const uint32_t smask = (1u << (so1 & 31u)) | (1u << (so2 & 31u));  // !XW: bit d set <=> syncmer

    __syncthreads();  // table + misc ready; from here on every warp works on its own

    // Software pipeline over tiles: the look-back + emission of tile A runs after the main loop
    // of the next tile B, so A's predecessors have published their counts by then.
    uint32_t p_valid = 0, p_tile = 0, p_inc = 0, p_spilled = 0, cur = 0;
    for (;;) {  // persistent: one tile (32 threads x S windows) per iteration, per warp
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(a.ticket, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        const bool have = tile < a.num_tiles;
        uint32_t cnt = 0, inc = 0, tile_spilled = 0;
        if (have) {
        const FSeg sg = make_fseg(a, tile, lane);
        const uint32_t qa0 = q0s + cur * QROWS * (uint32_t)ROWB;   // this lane's first row (shared address)
        const uint32_t qtrig = qa0 + a.q_trig * (uint32_t)ROWB;
        QT* const spill = reinterpret_cast<QT*>(sc0 + (size_t)cur * spill_words) + (size_t)lane * SPILLCAP;
        uint32_t qa = qa0, spilled = 0;
        {
            uint32_t fw = misc[1], rc = misc[2];
            // last valid window-end element + 1 (the first one is `lead` for every thread)
            const uint32_t e_hi = lead + sg.nvalid;

            // ---- thread-local view of the packed stream (32-bit word offsets from w0abs) -----
            // Entering bases of block b start at local base k-1+bW; leaving bases at local base
            // bW-1 (base -1 = virtual 'A').  Both are addressed with one extra virtual word in
            // front (+32 bits) so that the bit position never goes negative.
            const int64_t w0abs = sg.bit0 >> 5;
            // (kept in registers: ptxas otherwise re-derives it from a.seq in front of every prefetch)
            const uint32_t* const wbase = pinned(a.seq + w0abs);   // only dereferenced inside the buffer
            const uint32_t sh0 = (uint32_t)sg.bit0 & 31u;
            // may any (pre)fetch of this thread touch a word outside the buffer?
            const uint32_t wmax = ((sh0 + 32u + 2u * (k - 1) + 2u * (NB + 1) * SB) >> 5) + 3u;
            const bool clampd = w0abs < 1 || (uint64_t)w0abs + wmax >= a.seq_nwords;
            auto ldw = [&](uint32_t wl1) -> uint32_t {  // wl1 = word offset + 1 (virtual word 0)
                if (wl1 == 0) return 0u;
                if (clampd) return ld_word_any(a.seq, a.seq_nwords, (sg.bit0 >> 5) + (int64_t)wl1 - 1);
                return __ldg(wbase + (wl1 - 1));
            };
            // ---- prologue: consume k-1 bases, two per table step (leaving bases = virtual 'A') --
            {
                uint32_t pp = sh0 + 32u, rem = k - 1;
                while (rem) {
                    const uint32_t wl1 = pp >> 5, sh = pp & 31u;
                    uint32_t x = __funnelshift_r(ldw(wl1), ldw(wl1 + 1), sh);
                    uint32_t take = min(rem, 16u);
                    rem -= take;
                    pp += 32u;
                    for (; take >= 2; take -= 2, x >>= 4) {
                        const uint4 e = table(x & 15u);
                        fw = rotl32(fw, R2) ^ e.y;
                        if (HC) rc = rotr32(rc, R2) ^ e.w;
                    }
                    if (take) {
                        const uint4 e = table(x & 3u);
                        fw = rotl32(fw, R) ^ e.x;
                        if (HC) rc = rotr32(rc, R) ^ e.z;
                    }
                }
            }
            uint32_t pin = sh0 + 32u + 2u * (k - 1);  // bit position (+32) of the entering stream
            uint32_t pout = sh0 + 30u;                // bit position (+32) of the leaving stream
            uint32_t iw0 = ldw(pin >> 5), iw1 = ldw((pin >> 5) + 1), iw2 = SB > 16 ? ldw((pin >> 5) + 2) : 0u;
            uint32_t ow0 = ldw(pout >> 5), ow1 = ldw((pout >> 5) + 1), ow2 = SB > 16 ? ldw((pout >> 5) + 2) : 0u;
            ow0 &= ~(3u << (pout & 31u));  // element 0 leaves the virtual 'A'

            uint32_t RL[W], RR[W];
#pragma unroll
            for (int t = 0; t < W; t++) RL[t] = 0xffffffffu, RR[t] = 0u;
            uint32_t prev = 0xffffffffu;
            // XW: window of Wr k-mers ending at e = min over the sub-windows (W k-mers) ending at
            // e, e - W, ..., e - (T-1) W and e - (Wr - W); consecutive ones overlap or abut
            const uint32_t Dmax = Wr - W, T = XW ? (Wr + W - 1) / W - 1 : 0u;
            const uint32_t rmask = XW ? fast_ring_rows(Wr) - 1u : 0u;
            uint32_t tL[4] = {0, 0, 0, 0}, tR[4] = {0, 0, 0, 0};
            // ambiguity: only threads whose stretch holds an ambiguous base compute clean masks
            bool amb_here = false;
            uint32_t zrun = 0, pclean = 1;
            int64_t ab0 = 0;
            if (AMB) {
                // mask bit of local base 0 (bases in front of the sequence count as unambiguous)
                ab0 = ((sg.bit0 - a.bitbias) >> 1) + a.amb_bitbias;
                amb_here = amb_any(a, ab0, NB * SB + k - 1);
                if (amb_here) zrun = amb_clean_run(a, ab0, k - 1);
            }
            // the iteration after which the lead-in ends (B == 1); B > 1: after block 0 of iteration 0
            const uint32_t lead_b = lead / (uint32_t)SB - (B == 1 ? 1u : 0u);

            // Tiles at the end of a launch (and reads shorter than the longest of their tile): a lane
            // keeps computing after its last window -- on whatever follows, or on the clamped last word
            // of the buffer, a homopolymer that emits at every window -- and everything it pushes there
            // is garbage.  The rows of iterations that lie entirely behind the lane's last window are
            // dropped at once (the last partly valid iteration is trimmed after the loop), so the garbage
            // never fills the queue, let alone spills.
            const bool tile_partial = __any_sync(0xffffffffu, sg.nvalid < a.S);
            uint32_t qkeep = qa0;
            // a lane close to the end of its rows: the warp moves all its queues to the spill area
            auto spill_check = [&](uint32_t b) {
                if (__builtin_expect(__any_sync(0xffffffffu, qa > qtrig), 0)) {
                    if (B == 1 && b <= lead_b) {
                        qa = qa0;  // still inside the lead-in: nothing pushed so far is kept anyway
                    } else {
                        const uint32_t n = (qa - qa0) / (uint32_t)ROWB;
                        q_spill<ROWB, XW, QT>(qa0, n, spill + spilled);
                        spilled += n;
                        qa = qa0;
                        qkeep = qa0;
                        tile_spilled = 1;
                    }
                }
            };
            for (uint32_t b = 0; b < NB; b++) {
                const uint32_t eb = b * SB;
                const uint32_t shi = pin & 31u, sho = pout & 31u;
                const uint32_t x0 = __funnelshift_r(iw0, iw1, shi), x1 = SB > 16 ? __funnelshift_r(iw1, iw2, shi) : 0u;
                const uint32_t y0 = __funnelshift_r(ow0, ow1, sho), y1 = SB > 16 ? __funnelshift_r(ow1, ow2, sho) : 0u;
                // prefetch the next block's words (consumed ~W*30 instructions later)
                pin += 2u * SB;
                pout += 2u * SB;
                if (!clampd) {
                    const uint32_t* pi = wbase + (pin >> 5) - 1;
                    const uint32_t* po = wbase + (pout >> 5) - 1;
                    iw0 = __ldg(pi), iw1 = __ldg(pi + 1);
                    ow0 = __ldg(po), ow1 = __ldg(po + 1);
                    if (SB > 16) iw2 = __ldg(pi + 2), ow2 = __ldg(po + 2);
                } else {
                    iw0 = ldw(pin >> 5), iw1 = ldw((pin >> 5) + 1);
                    ow0 = ldw(pout >> 5), ow1 = ldw((pout >> 5) + 1);
                    if (SB > 16) iw2 = ldw((pin >> 5) + 2), ow2 = ldw((pout >> 5) + 2);
                }
                uint32_t preL = 0, preR = 0;
                uint32_t gr[4] = {0, 0, 0, 0}, gm[4] = {0, 0, 0, 0}, gp[4] = {0, 0, 0, 0};  // group of four windows
                uint32_t clean = 0xffffffffu;
                if (AMB && amb_here) {
                    // window ending at k-mer eb+t = the l bases ending at local base eb+t+k-1.
                    // A clean window right after an ambiguous one is always emitted: the
                    // reference compares against SKIPPED there (src/intrinsics/dedup.rs:147-155).
                    const uint2 cm = amb_clean_mask(a, ab0 + eb + (k - 1), SB, a.l, zrun);
                    clean = cm.x;
                    zrun = cm.y;
                }
#pragma unroll
                for (int j = 0; j < B; j++) {  // van-Herk block j of this iteration
                const int o = j * W;           // its first k-mer inside the iteration
                // this block's 2W bits of both streams, interleaved: byte = (in,in,out,out)
                uint32_t xs0 = x0, xs1 = x1, ys0 = y0, ys1 = y1;
                if (o != 0) {  // B > 1 implies W <= 16: one 32-bit word per stream is enough
                    xs0 = 2 * o < 32 ? __funnelshift_r(x0, x1, 2 * o) : (x1 >> ((2 * o - 32) & 31));
                    ys0 = 2 * o < 32 ? __funnelshift_r(y0, y1, 2 * o) : (y1 >> ((2 * o - 32) & 31));
                }
                uint32_t N[4];
                N[0] = (xs0 & 0x0F0F0F0Fu) | ((ys0 << 4) & 0xF0F0F0F0u);
                N[1] = ((xs0 >> 4) & 0x0F0F0F0Fu) | (ys0 & 0xF0F0F0F0u);
                if (W > 16) {
                    N[2] = (xs1 & 0x0F0F0F0Fu) | ((ys1 << 4) & 0xF0F0F0F0u);
                    N[3] = ((xs1 >> 4) & 0x0F0F0F0Fu) | (ys1 & 0xF0F0F0F0u);
                }
                // Two k-mers per step: one table load gives both hashes; the prefix minimum and
                // the first window of the pair use the 3-input VIMNMX3.
                constexpr int TD = MZ_FAST_TD;  // table lookups in flight (pairs ahead)
                uint4 tq[TD];
#pragma unroll
                for (int d = 0; d < TD; d++) tq[d] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                for (int t = 0; t < W; t += 2) {
                    const bool two = t + 1 < W;
                    if (XW && (t & 3) == 0) {
                        // taps of k-mers t .. t+3: sub-window minima that ended i*W (i < T) and
                        // Wr - W k-mers earlier; they were written >= 16 k-mers ago
#pragma unroll
                        for (int u = 0; u < 4; u++) tL[u] = 0xffffffffu, tR[u] = 0u;
                        for (uint32_t i = 1; i <= T; i++) {
                            const uint32_t d = i < T ? i * (uint32_t)W : Dmax;
                            const uint32_t sbase = eb + o + t - d;
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                if (t + u < W) {
                                    const uint32_t sl = (sbase + u) & rmask;
                                    if (LR) {
                                        const uint2 v = (reinterpret_cast<const uint2*>(r1) + lane)[(size_t)sl * 32];
                                        tL[u] = min(tL[u], v.x), tR[u] = max(tR[u], v.y);
                                    } else {
                                        tL[u] = min(tL[u], r1[(size_t)sl * 32 + lane]);
                                    }
                                }
                            }
                        }
                    }
                    // table entry of this pair (requested one pair ahead: the lookup only depends on
                    // the block's base bytes, and its latency is the main loop's largest stall)
                    auto lookup = [&](int tt) -> uint4 {
                        const uint32_t word = N[(tt >> 4) * 2 + ((tt >> 1) & 1)];
                        const uint32_t idx = get_byte(word, (tt & 15) >> 2);
                        const uint32_t addr = idx * (16u * FAST_TC) + tb;
                        if (HC) return lds128(addr);
                        const uint2 e2 = lds64(addr);
                        return make_uint4(e2.x, e2.y, 0u, 0u);
                    };
                    // (instances that have no registers to spare for the queue look up just in time: long
                    // windows, w = 201 60 -> 53 Gbp/s with the queue; strand-aware kernels with two suffix
                    // arrays of more than 26 registers each, k = 11 w = 31 463 -> 450)
                    uint4 e;
                    if (XW || (LR && W > 26)) {
                        e = lookup(t);
                    } else {
                        if (t == 0) {
#pragma unroll
                            for (int d = 0; d < TD; d++)
                                if (2 * d < W) tq[d] = lookup(2 * d);
                        }
                        e = tq[0];
#pragma unroll
                        for (int d = 0; d + 1 < TD; d++) tq[d] = tq[d + 1];
                        if (t + 2 * TD < W) tq[TD - 1] = lookup(t + 2 * TD);
                    }
                    uint32_t h0, h1 = 0;
                    if (HC) {
                        const uint32_t fA = rotl32(fw, R) ^ e.x, rA = rotr32(rc, R) ^ e.z;
                        h0 = fA + rA;
                        if (two) {
                            fw = rotl32(fw, R2) ^ e.y;
                            rc = rotr32(rc, R2) ^ e.w;
                            h1 = fw + rc;
                        } else {
                            fw = fA;
                            rc = rA;
                        }
                    } else {
                        const uint32_t fA = rotl32(fw, R) ^ e.x;
                        h0 = fA;
                        if (two) {
                            fw = rotl32(fw, R2) ^ e.y;
                            h1 = fw;
                        } else {
                            fw = fA;
                        }
                    }
                    const uint32_t pos0 = imad(eb, one, o + t), pos1 = imad(eb, one, o + t + 1);
                    const uint32_t le0 = (h0 & 0xffff0000u) | pos0, le1 = (h1 & 0xffff0000u) | pos1;
                    uint32_t res0, res1 = 0, mR0 = 0, mR1 = 0;
                    {
                        const uint32_t s1 = t + 1 < W ? RL[t + 1 < W ? t + 1 : 0] : 0u;
                        const uint32_t s2 = t + 2 < W ? RL[t + 2 < W ? t + 2 : 0] : 0u;
                        if (t == 0) {
                            res0 = W > 1 ? min(le0, s1) : le0;
                            preL = two ? min(le0, le1) : le0;
                        } else {
                            res0 = t + 1 < W ? __vimin3_u32(preL, le0, s1) : min(preL, le0);
                            preL = two ? __vimin3_u32(preL, le0, le1) : min(preL, le0);
                        }
                        if (two) res1 = t + 2 < W ? min(preL, s2) : preL;
                        RL[t] = le0;
                        if (two) RL[t + 1] = le1;
                    }
                    if (LR) {
                        const uint32_t re0 = le0 ^ 0xffff0000u, re1 = le1 ^ 0xffff0000u;
                        const uint32_t s1 = t + 1 < W ? RR[t + 1 < W ? t + 1 : 0] : 0u;
                        const uint32_t s2 = t + 2 < W ? RR[t + 2 < W ? t + 2 : 0] : 0u;
                        if (t == 0) {
                            mR0 = W > 1 ? max(re0, s1) : re0;
                            preR = two ? max(re0, re1) : re0;
                        } else {
                            mR0 = t + 1 < W ? __vimax3_u32(preR, re0, s1) : max(preR, re0);
                            preR = two ? __vimax3_u32(preR, re0, re1) : max(preR, re0);
                        }
                        if (two) mR1 = t + 2 < W ? max(preR, s2) : preR;
                        RR[t] = re0;
                        if (two) RR[t + 1] = re1;
                    }
                    if (XW) {
                        // level 1 -> ring (read back Dmax .. k-mers later), level 2 = minimum with
                        // the taps fetched at the head of this group of four
                        const uint32_t s0 = (eb + o + t) & rmask, s1 = (eb + o + t + 1) & rmask;
                        if (LR) {
                            uint2* const row = reinterpret_cast<uint2*>(r1) + lane;
                            row[(size_t)s0 * 32] = make_uint2(res0, mR0);
                            if (two) row[(size_t)s1 * 32] = make_uint2(res1, mR1);
                            mR0 = max(mR0, tR[t & 3]);
                            if (two) mR1 = max(mR1, tR[(t + 1) & 3]);
                        } else {
                            r1[(size_t)s0 * 32 + lane] = res0;
                            if (two) r1[(size_t)s1 * 32 + lane] = res1;
                        }
                        res0 = min(res0, tL[t & 3]);
                        if (two) res1 = min(res1, tL[(t + 1) & 3]);
                    }
                    // ---- group of four windows: tie test, then one queue entry per new selection ----
                    gr[t & 2] = res0, gm[t & 2] = mR0, gp[t & 2] = pos0;
                    if (two) gr[(t & 2) + 1] = res1, gm[(t & 2) + 1] = mR1, gp[(t & 2) + 1] = pos1;
                    if ((t & 2) || t + 2 >= W) {
                        const int g0 = t & ~3;                    // first window of the group (inside the block)
                        const int ng = W - g0 < 4 ? W - g0 : 4;   // windows in it
                        if (LR) {
                            // leftmost != rightmost in one of them?  res ^ mR has all sixteen key bits
                            // set and the xor of the two positions below them.  Cold: strand rule.
                            uint32_t tie = gr[0] ^ gm[0];
#pragma unroll
                            for (int u = 1; u < 4; u++)
                                if (u < ng) tie |= gr[u] ^ gm[u];
                            if (__builtin_expect(tie != 0xffff0000u, 0)) {
                                const uint4 f = tie_fix4(a, sg.bit0, gp[0], (uint32_t)ng, make_uint4(gr[0], gr[1], gr[2], gr[3]),
                                                         make_uint4(gm[0], gm[1], gm[2], gm[3]));
                                gr[0] = f.x, gr[1] = f.y, gr[2] = f.z, gr[3] = f.w;
                            }
                        }
                        if (!SYNC && !AMB) {
                            uint32_t ge[4];
#pragma unroll
                            for (int u = 0; u < 4; u++)
                                ge[u] = u < ng ? (XW ? imad(gr[u] & 0xffffu, DMUL, gp[u]) : imad(gr[u], DMUL, gp[u])) : 0u;
                            if (ng == 4) q_push4_ne_fma<XW, 4>(qa, prev, gr, ge, one);
                            else if (ng == 3) q_push4_ne_fma<XW, 3>(qa, prev, gr, ge, one);
                            else if (ng == 2) q_push4_ne_fma<XW, 2>(qa, prev, gr, ge, one);
                            else q_push4_ne_fma<XW, 1>(qa, prev, gr, ge, one);
                            prev = gr[ng - 1];
                        } else if (SYNC && !AMB && !XW) {
                            uint32_t ge[4], gk[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                ge[u] = u < ng ? imad(gr[u], DMUL, gp[u]) : 0u;
                                gk[u] = u < ng ? smask : 0u;
                            }
                            q_push4_mask_fma(qa, gk, ge, one);
                        } else {
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                if (u >= ng) continue;
                                const int tt = o + g0 + u;            // window inside the iteration
                                const uint32_t r = gr[u], ps = gp[u];
                                const uint32_t ent = XW ? imad(r & 0xffffu, DMUL, ps) : imad(r, DMUL, ps);
                                bool pe;
                                if (SYNC) {
                                    const uint32_t d = ps - (r & 0xffffu);
                                    pe = d == so1 || d == so2;
                                } else {
                                    pe = r != prev;
                                    prev = r;
                                }
                                if (AMB) {
                                    const bool c = (clean >> tt) & 1u;
                                    const bool cm1 = tt ? ((clean >> ((tt ? tt : 1) - 1)) & 1u) : (pclean != 0u);
                                    if (!SYNC) pe = pe || !cm1;
                                    pe = pe && c;
                                }
                                q_push_if<ROWB, XW>(qa, pe, ent);
                            }
                        }
                        if (XW) spill_check(b);
                    }
                }
                // suffix minima of this block (slot 0 is never needed)
#pragma unroll
                for (int q = W - 2; q >= 1; q--) {
                    RL[q] = min(RL[q], RL[q + 1]);
                    if (LR) RR[q] = max(RR[q], RR[q + 1]);
                }
                // end of the lead-in: everything pushed so far belongs to windows in front of the
                // segment; the last of them seeds the dedup comparison unless nothing is in front
                if ((B > 1 && j == 0) || (B == 1 && j == B - 1)) {
                    if (b == lead_b) {
                        qa = qa0;
                        spilled = 0;
                        if (sg.first_always) prev = 0xffffffffu;
                    }
                }
                }  // van-Herk blocks of this iteration
                if (AMB) pclean = (clean >> (SB - 1)) & 1u;
                if (tile_partial) {
                    if (eb >= e_hi) qa = qkeep;
                    else qkeep = qa;
                    // every lane is behind its last window: the rest of the tile is idle work
                    if (__all_sync(0xffffffffu, eb + (uint32_t)SB >= e_hi)) break;
                }
                if (!XW) spill_check(b);
            }
            // ---- count: rows pushed, minus those of windows behind the segment's last one ------
            uint32_t n = (qa - qa0) / (uint32_t)ROWB;
            if (tile_spilled) {  // warp-uniform: everything goes to the spill area
                q_spill<ROWB, XW, QT>(qa0, n, spill + spilled);
                n += spilled;
            }
            if (__any_sync(0xffffffffu, e_hi != NB * (uint32_t)SB)) {
                if (sg.nvalid == 0) n = 0;
                while (n) {
                    const uint32_t v = tile_spilled ? (uint32_t)spill[n - 1] : q_load<XW>(qa0 + (n - 1) * (uint32_t)ROWB);
                    if ((v >> DBITS) + (v & DMUL) < e_hi) break;
                    n--;
                }
            }
            cnt = n;
        }
        // publish this tile's count (aggregate) right away; its prefix is resolved one tile later
        __syncwarp();
        inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += v;
        }
        if (lane == 31) st_state(a.tile_state + tile, ((unsigned long long)inc << 2) | 1ull);
        uint32_t* const tp = tp0 + cur * FAST_TOFFS;
        tp[lane] = inc - cnt;
        if (lane == 31) tp[32] = inc;
        }  // have

        // ---- ordered emission of the PREVIOUS tile (of this one without a second queue buffer),
        //      warp-autonomous (no block barriers) ---------------------------------------------------
        if (!defer) p_valid = have ? 1u : 0u, p_tile = tile, p_inc = inc, p_spilled = tile_spilled;
        if (p_valid) {
        const uint32_t tile_e = p_tile, inc_e = p_inc, pb = defer ? (cur ^ 1u) : 0u;
        const uint32_t total = __shfl_sync(0xffffffffu, inc_e, 31);
        const unsigned long long gbase = lookback_excl(a.tile_state, tile_e, total);
        if (lane == 0 && tile_e == a.num_tiles - 1) {
            *a.count_out = gbase + total;
            if (a.h_count) *a.h_count = gbase + total;
        }
        const bool ovf = gbase + total > a.cap;
        if (ovf && lane == 0) {
            *a.overflow = 1u;
            if (a.h_overflow) *a.h_overflow = 1u;
        }
        if (a.n_reads != 0) write_csr_offset(a, (uint64_t)tile_e * 32u + lane, gbase + inc_e);
        if (!ovf && total != 0) {
            __syncwarp();
            const uint32_t* const tpe = tp0 + pb * FAST_TOFFS;
            const unsigned char* ebase;
            uint32_t sI, sT;
            if (p_spilled) {
                ebase = reinterpret_cast<const unsigned char*>(sc0 + (size_t)pb * spill_words);
                sI = (uint32_t)sizeof(QT), sT = SPILLCAP * (uint32_t)sizeof(QT);
                __threadfence_block();
            } else {
                ebase = q0 + (size_t)pb * QROWS * ROWB;
                sI = (uint32_t)ROWB, sT = (uint32_t)sizeof(QT);
            }
            if (a.n_reads == 0 && !p_spilled && !(a.value_bits != 0 && fast_tile_at_buffer_end(a, tile_e))) {
                const uint32_t qs = q0s - lane * (uint32_t)sizeof(QT) + pb * QROWS * (uint32_t)ROWB;
                fast_emit_seq<XW, SYNC, LR>(a, tile_e, gbase, total, (uint32_t)__cvta_generic_to_shared(tpe), qs);
            } else {
                if (a.value_bits == 64) fast_emit_generic<64, XW>(a, tile_e, gbase, total, tpe, ebase, sI, sT);
                else if (a.value_bits == 128) fast_emit_generic<128, XW>(a, tile_e, gbase, total, tpe, ebase, sI, sT);
                else fast_emit_generic<0, XW>(a, tile_e, gbase, total, tpe, ebase, sI, sT);
            }
            __syncwarp();
        }
        }  // p_valid
        if (!have) break;
        if (defer) {
            p_valid = 1, p_tile = tile, p_inc = inc, p_spilled = tile_spilled;
            cur ^= 1u;
        }
    }
}

// ---- host side ---------------------------------------------------------------------------
struct FastPlan {
    uint32_t S = 0, num_tiles = 0, grid = 0, q_rows = 0, q_trig = 0, q_bufs = 2, nb = 0, lead = 0;
    size_t scratch_words_per_block = 0;  // spill area, per warp and buffer
    size_t r1_words = 0;                 // XW: level-1 rows, per warp
};

// queue geometry for segments of S windows; false when it does not fit the shared memory
inline bool fast_queue_plan(uint32_t S, const mz_params& p, FastPlan* pl, bool keep_l1 = false) {
    const uint32_t sb = fast_sb(fast_wt(p.w));
    pl->S = S;
    pl->lead = fast_lead(p.w);
    pl->nb = fast_nb(S, p.w);
    pl->q_trig = fast_q_trig(S, p);
    // rows pushed between two overflow checks: an iteration (one check at its end), long windows
    // (4-byte entries, sparse output) check after every group of four windows instead
    pl->q_rows = pl->q_trig + (p.w > FAST_MAX_W ? 4u : sb);
    pl->scratch_words_per_block = fast_spill_words(S, p.w);
    pl->r1_words = fast_r1_words(p.w, p.strand_tiebreak != 0);
    // entry field widths: selected k-mer below 2^11 (u16 entries) / 2^16 (positions in the keys)
    const uint32_t elems = pl->nb * sb + 1;
    if (p.w <= FAST_MAX_W ? elems >= 2048 : elems >= 65535) return false;
    pl->q_bufs = fast_q_bufs(p);
    return fast_smem(p.w, pl->q_rows, pl->q_bufs) <= fast_smem_limit(keep_l1);
}

// Geometry for the fast kernel; returns false when (k, w, ...) is outside its domain.
inline bool plan_fast(int sm_count, const mz_params& p, uint64_t nwin, FastPlan* pl, bool allow_xw = false) {
    const bool xw = p.w > FAST_MAX_W;
    if (xw && (!allow_xw || p.w > FAST_XW_MAX_W || getenv("MZ_NO_XW"))) return false;
    const char* env_s = getenv("MZ_FAST_S");
    const char* env_bps = getenv("MZ_FAST_BPS");
    const uint32_t bps = env_bps ? (uint32_t)atoi(env_bps) : FAST_BPS;  // resident blocks per SM
    const uint64_t slots = (uint64_t)sm_count * bps * FAST_WARPS;  // resident warps
    const uint32_t sb = fast_sb(fast_wt(p.w)), lead = fast_lead(p.w);
    uint32_t s;
    if (env_s) {
        s = (uint32_t)atoi(env_s);
    } else {
        // long segments amortise the (k+w-2)-base warm-up.  Few waves (small inputs, shards of a
        // multi-GPU run): equal tiles in a whole number of waves, m = 1 .. 8 tiles per warp.
        // (Extending this to m <= 48 -- an eighth of C2 is 12.2 tiles per warp -- was measured:
        // 0.641 vs 0.644 ms, the ticket counter already evens the tail out.)
        static const uint32_t smax0 = getenv("MZ_FAST_SMAX") ? (uint32_t)atoi(getenv("MZ_FAST_SMAX")) : 420u;
        const uint32_t smax = xw ? smax0 + p.w : smax0;
        const uint64_t per_wave = slots * 32;
        const uint64_t m = (nwin + per_wave * smax - 1) / (per_wave * smax);
        const uint64_t want = m <= 8 ? (nwin + per_wave * m - 1) / (per_wave * m) : smax;
        s = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(want, 64), smax);
        // a thread computes lead + S k-mers in whole loop iterations of SB k-mers: pick S so that
        // the last iteration is full (S = NB*SB - lead)
        uint32_t nb = std::max<uint32_t>(1, (s + lead + (m <= 8 ? sb - 1 : 0)) / sb);  // few waves: round up
        while (nb * sb < lead + 16) nb++;
        s = nb * sb - lead;
    }
    s = std::max<uint32_t>(1, s);
    // the queues live in shared memory: shorten the segments until they fit (dense outputs).
    // First within the budget that keeps 32 KB of L1; when that costs more than a tenth of the
    // segment length, with all the shared memory there is.
    auto fit = [&](uint32_t s0, bool keep_l1, FastPlan* out) -> uint32_t {
        uint32_t q = s0;
        while (!fast_queue_plan(q, p, out, keep_l1)) {
            if (q <= sb) {
                if (q <= 1) return 0;
                q = std::max<uint32_t>(1, q / 2);
            } else {
                q -= sb;
            }
        }
        return q;
    };
    const uint32_t s_want = s;
    s = fit(s_want, true, pl);
    if (s == 0 || (uint64_t)s * 10 < (uint64_t)s_want * 9) {
        FastPlan alt = *pl;
        const uint32_t s2 = fit(s_want, false, &alt);
        if (s2 == 0 && s == 0) return false;
        if (s2 > s) {
            s = s2;
            *pl = alt;
        }
    }
    const uint64_t Tt = (uint64_t)32 * s;
    const uint64_t tiles = (nwin + Tt - 1) / Tt;
    if (tiles == 0 || tiles > 0x7fffffffull) return false;
    pl->num_tiles = (uint32_t)tiles;
    // one block per SM as soon as there are that many tiles (warps take tiles from the ticket
    // counter, so a small launch spreads over all SMs instead of filling a few of them)
    pl->grid = (uint32_t)std::min<uint64_t>(tiles, (uint64_t)sm_count * bps);
    return true;
}

// copy a plan's queue geometry into the kernel arguments
inline void fast_plan_args(const FastPlan& fp, KArgs& a) {
    a.S = fp.S;
    a.q_rows = fp.q_rows;
    a.q_bufs = fp.q_bufs;
    a.q_trig = fp.q_trig;
    a.nb = fp.nb;
    a.lead = fp.lead;
    a.one = 1;
    a.scratch_words_per_block = fp.scratch_words_per_block;
    a.r1_words_per_warp = fp.r1_words;
}

template <int W, bool HC, bool LR, bool SYNC, bool AMB = false, bool XW = false>
inline int launch_fast_inst(uint32_t grid, const KArgs& a, cudaStream_t st) {
    auto kern = mz_fast_kernel<W, HC, LR, SYNC, AMB, XW>;
    const size_t smem = fast_smem(a.w, a.q_rows, a.q_bufs);
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return MZ_ERR_CUDA;
    kern<<<grid, FAST_NT, smem, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? MZ_OK : MZ_ERR_CUDA;
}

template <int W>
inline int launch_fast_w(const mz_params& p, uint32_t grid, const KArgs& a, cudaStream_t st) {
    const bool sync = p.mode != MZ_MODE_MINIMIZER;
    if (p.strand_tiebreak) {
        return sync ? launch_fast_inst<W, true, true, true>(grid, a, st)
                    : launch_fast_inst<W, true, true, false>(grid, a, st);
    }
    if (p.hash_canonical) {  // forward builder + canonical hasher (src/minimizers.rs:69-71)
        return sync ? launch_fast_inst<W, true, false, true>(grid, a, st)
                    : launch_fast_inst<W, true, false, false>(grid, a, st);
    }
    return sync ? launch_fast_inst<W, false, false, true>(grid, a, st)
                : launch_fast_inst<W, false, false, false>(grid, a, st);
}

// skip-ambiguous instances: canonical builders only (src/lib.rs:451)
template <int W>
inline int launch_fast_amb_w(const mz_params& p, uint32_t grid, const KArgs& a, cudaStream_t st) {
    if (!p.strand_tiebreak) return MZ_ERR_NOT_CANONICAL;
    return p.mode != MZ_MODE_MINIMIZER ? launch_fast_inst<W, true, true, true, true>(grid, a, st)
                                       : launch_fast_inst<W, true, true, false, true>(grid, a, st);
}

// defined in mz_fast_g{0..3}.cu (W = 1..8, 9..16, 17..24, 25..32)
int launch_fast_g0(const mz_params&, uint32_t, const KArgs&, cudaStream_t);
int launch_fast_g1(const mz_params&, uint32_t, const KArgs&, cudaStream_t);
int launch_fast_g2(const mz_params&, uint32_t, const KArgs&, cudaStream_t);
int launch_fast_g3(const mz_params&, uint32_t, const KArgs&, cudaStream_t);

// long windows (XW): instance of the sub-window length W = fast_wt(p.w) in 17..24
template <int W>
inline int launch_fast_xw_w(const mz_params& p, uint32_t grid, const KArgs& a, cudaStream_t st) {
    const bool sync = p.mode != MZ_MODE_MINIMIZER;
    if (a.amb) {  // skip-ambiguous: canonical builders only (src/lib.rs:451)
        if (!p.strand_tiebreak) return MZ_ERR_NOT_CANONICAL;
        return sync ? launch_fast_inst<W, true, true, true, true, true>(grid, a, st)
                    : launch_fast_inst<W, true, true, false, true, true>(grid, a, st);
    }
    if (p.strand_tiebreak) {
        return sync ? launch_fast_inst<W, true, true, true, false, true>(grid, a, st)
                    : launch_fast_inst<W, true, true, false, false, true>(grid, a, st);
    }
    if (p.hash_canonical) {
        return sync ? launch_fast_inst<W, true, false, true, false, true>(grid, a, st)
                    : launch_fast_inst<W, true, false, false, false, true>(grid, a, st);
    }
    return sync ? launch_fast_inst<W, false, false, true, false, true>(grid, a, st)
                : launch_fast_inst<W, false, false, false, false, true>(grid, a, st);
}
// defined in mz_fast_x{0,1}.cu (sub-window 17..20, 21..24)
int launch_fast_x0(const mz_params&, uint32_t, const KArgs&, cudaStream_t);
int launch_fast_x1(const mz_params&, uint32_t, const KArgs&, cudaStream_t);

// defined in mz_fast_a{0..3}.cu
int launch_fast_a0(const mz_params&, uint32_t, const KArgs&, cudaStream_t);
int launch_fast_a1(const mz_params&, uint32_t, const KArgs&, cudaStream_t);
int launch_fast_a2(const mz_params&, uint32_t, const KArgs&, cudaStream_t);
int launch_fast_a3(const mz_params&, uint32_t, const KArgs&, cudaStream_t);

inline int launch_fast(const mz_params& p, uint32_t grid, const KArgs& a, cudaStream_t st) {
    if (p.w > FAST_MAX_W) {
        return fast_wt(p.w) <= 20 ? launch_fast_x0(p, grid, a, st) : launch_fast_x1(p, grid, a, st);
    }
    if (a.amb) {
        switch ((p.w - 1) / 8) {
            case 0: return launch_fast_a0(p, grid, a, st);
            case 1: return launch_fast_a1(p, grid, a, st);
            case 2: return launch_fast_a2(p, grid, a, st);
            case 3: return launch_fast_a3(p, grid, a, st);
            default: return MZ_ERR_UNSUPPORTED;
        }
    }
    switch ((p.w - 1) / 8) {
        case 0: return launch_fast_g0(p, grid, a, st);
        case 1: return launch_fast_g1(p, grid, a, st);
        case 2: return launch_fast_g2(p, grid, a, st);
        case 3: return launch_fast_g3(p, grid, a, st);
        default: return MZ_ERR_UNSUPPORTED;
    }
}

}  // namespace mz
