// mz_fast.cuh -- W-specialised, register-resident minimizer / syncmer kernel for sm_100a.
//
// Persistent warps pull tiles (32 threads x S windows) from a ticket counter; every warp is an
// autonomous worker (own look-back, own staging), so there is no block barrier after start-up.
// One thread walks S consecutive windows.  Per van-Herk block of W k-mers (fully unrolled):
//   * the words holding the entering- and leaving-base streams were prefetched one block ahead;
//     they are re-aligned with funnel shifts and interleaved so that every byte holds
//     (in,in,out,out) of two consecutive bases;
//   * one LDS.128 from a 256-entry table returns the rolling-hash deltas of BOTH bases for the
//     forward and the reverse-complement hash (ntHash/mulHash are GF(2)-linear in the tables),
//     so a k-mer hash costs SHF+LOP3 per strand;
//   * (hash & 0xffff0000) | pos goes through a prefix-min / suffix-min pair whose W-entry suffix
//     array lives in registers (static indexing); the rightmost minimum uses max on the
//     complemented key, exactly the reference's packing (src/sliding_min.rs:190-195,336-341);
//   * leftmost != rightmost is detected once per iteration from the packed position bytes of
//     the two chains (one LOP3 per four windows); the strand rule
//     (src/canonical.rs) runs in a cold fix-up for the few blocks that contain such a window;
//   * the low byte of the selected position and a flag bit per window go to an L2-resident
//     global scratch (coalesced rows); the warp then turns them into ordered, coalesced output
//     (look-back over tile descriptors, staging list, software-pipelined emission pass).
#pragma once
#include "../../include/mz_b200.h"
#include <algorithm>
#include <cstdlib>
#include "mz_emit.cuh"

namespace mz {

constexpr uint32_t FAST_MAX_W = 32;
// One block of 16 autonomous warps per SM.  (Warps never synchronise after start-up, so the block
// size only decides how many warps share one copy of the hash table in shared memory.)
constexpr uint32_t FAST_NT = 512;   // threads per block
constexpr uint32_t FAST_BPS = 1;    // resident blocks per SM
// The 256-entry delta table is stored FAST_TC times, copy c = lane & 7 interleaved so that entry
// e of copy c sits at 16-byte slot e*8 + c: the eight lanes of a quarter warp (one LDS.128 pass)
// always hit eight different bank groups -> no bank conflicts on the random table lookups.
constexpr uint32_t FAST_TC = 8;
constexpr size_t FAST_SMEM_LIMIT = 194 * 1024;  // per block (196 KB carve-out): ~60 KB of L1 stay for the input stream

// Per-block scratch in GLOBAL memory (L2-resident: a few hundred KB per resident block, reused
// for every tile the persistent block processes).  Row r, thread t -> word scratch[r*NT + t]:
//   rows (b*WQ + q)   low bytes of the selected positions of windows 4q..4q+3 of van-Herk block b
// The flag word of block b (bit t <-> window ending at element bW+t) stays in shared memory.
// Small windows: several van-Herk blocks (B of them, B*W <= 32 k-mers) share one loop iteration, so
// the per-iteration ingest/record overhead is amortised over ~32 k-mers for every W.
__host__ __device__ constexpr uint32_t fast_b(uint32_t W) { return W <= 16 ? 32 / W : 1; }
__host__ __device__ constexpr uint32_t fast_sb(uint32_t W) { return fast_b(W) * W; }
__host__ __device__ constexpr uint32_t fast_wq(uint32_t W) { return (fast_sb(W) + 3) / 4; }
// Windows longer than FAST_MAX_W (XW instances): the minimum over w k-mers is the minimum over
// T+1 shifted sub-windows of wt = fast_wt(w) <= 24 k-mers each; the kernel instance is the one of
// wt.  For w <= FAST_MAX_W, wt = w.
constexpr uint32_t FAST_XW_MAX_W = 255;  // selected position - window start must fit a byte
__host__ __device__ inline uint32_t fast_wt(uint32_t w) {
    if (w <= FAST_MAX_W) return w;
    const uint32_t parts = (w + 23) / 24;
    return (w + parts - 1) / parts;
}
__host__ __device__ inline uint32_t fast_nb(uint32_t S, uint32_t w) {
    const uint32_t sb = fast_sb(fast_wt(w));
    return (S + 1 + (w - 1) + sb - 1) / sb;  // k-mers = S + has_prev + w - 1
}
// scratch words per WARP (a warp is an autonomous worker: tile = 32 threads x S windows)
inline size_t fast_scratch_words(uint32_t S, uint32_t w) {
    return (size_t)fast_nb(S, w) * fast_wq(fast_wt(w)) * 32;
}
// XW instances: words per warp for the ring of level-1 results (one row of 32 lanes per k-mer;
// leftmost and rightmost sub-window minimum for strand-aware builders)
__host__ __device__ inline uint32_t fast_ring_rows(uint32_t w) {  // power of two >= (w - wt) + iteration
    const uint32_t need = (w - fast_wt(w)) + fast_sb(fast_wt(w)) + 4;
    uint32_t r = 32;
    while (r < need) r <<= 1;
    return r;
}
inline size_t fast_r1_words(uint32_t S, uint32_t w, bool lr) {
    (void)S;
    return w <= FAST_MAX_W ? 0 : (size_t)fast_ring_rows(w) * 32 * (lr ? 2 : 1);
}
constexpr uint32_t FAST_WARPS = FAST_NT / 32;
// shared memory: table | misc | per-warp staging list | per-warp flag words (2 tiles in flight)
inline size_t fast_smem(uint32_t S, uint32_t W, uint32_t list_cap) {
    return 256 * FAST_TC * 16 + 32 + (size_t)FAST_WARPS * list_cap * 4 + (size_t)FAST_WARPS * 2 * fast_nb(S, W) * 32 * 4;
}
// staging entries per warp: 1.5x the expected emissions of a tile (a second pass handles more)
inline uint32_t fast_list_cap(uint32_t S, const mz_params& p) {
    const double dens = p.mode == MZ_MODE_MINIMIZER ? 2.0 / (p.w + 1.0)
                      : p.mode == MZ_MODE_CLOSED_SYNCMER ? (p.w == 1 ? 1.0 : 2.0 / p.w) : 1.0 / p.w;
    static const double slack = getenv("MZ_FAST_LISTF") ? atof(getenv("MZ_FAST_LISTF")) : 1.5;
    const uint32_t want = (uint32_t)(32.0 * S * dens * slack) + 64;
    // ... but never more than what keeps the block within FAST_SMEM_LIMIT: dense outputs
    // (small w) simply take more staging passes per tile
    const size_t fixed = fast_smem(S, p.w, 0);
    const uint32_t fit = fixed + 256 * 4 * FAST_WARPS >= FAST_SMEM_LIMIT
                             ? 256u : (uint32_t)((FAST_SMEM_LIMIT - fixed) / (4 * FAST_WARPS)) / 128 * 128;
    return std::max<uint32_t>(std::min<uint32_t>((want + 127) / 128 * 128, std::min<uint32_t>(fit, 4096)), 256);
}

// Rare path (leftmost != rightmost minimum): strand rule 2*#TG > l on the window's l bases
// (src/canonical.rs:19-29).  Kept out of line so the unrolled hot loop stays small.
static __device__ __noinline__ bool window_prefers_left(const uint32_t* wbase, uint32_t sh0, uint32_t wlim,
                                                 uint32_t lbit, uint32_t l) {
    uint32_t p = sh0 + lbit, bits = 2u * l, cnt = 0;
    while (bits) {
        const uint32_t wl = p >> 5, sh = p & 31u;
        const uint32_t w0 = __ldg(wbase + min(wl, wlim)), w1 = __ldg(wbase + min(wl + 1, wlim));
        uint32_t v = __funnelshift_r(w0, w1, sh) & 0xAAAAAAAAu;
        const uint32_t take = bits < 32u ? bits : 32u;
        if (take < 32u) v &= (1u << take) - 1u;
        cnt += __popc(v);
        p += take;
        bits -= take;
    }
    return 2u * cnt > l;
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
// bf |= bit when a != b (forced to SETP + predicated OR: two issue slots)
__device__ __forceinline__ void or_if_ne(uint32_t& bf, uint32_t a, uint32_t b, uint32_t bit) {
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(bf) : "r"(a), "r"(b), "r"(bit));
}
// a*b + c written as mad.lo in PTX, b = 1 from a volatile mov.  ptxas still picks the pipe per
// use, but with plain C++ adds for the positions it allocates 128 instead of 115 registers and
// the w = 31 instance runs 19 % slower (measured), so the positions stay expressed this way.
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
// J is a compile-time constant after unrolling
__device__ __forceinline__ uint32_t put_byte(uint32_t acc, uint32_t v, int J) {  // acc.byte[J] = v.byte[0]
    return __byte_perm(acc, v, J == 0 ? 0x3214 : J == 1 ? 0x3240 : J == 2 ? 0x3410 : 0x4210);
}
__device__ __forceinline__ uint32_t get_byte(uint32_t w, int J) {
    return J == 0 ? (w & 0xffu) : J == 3 ? (w >> 24) : __byte_perm(w, 0, 0x4440 + J);
}

// AMB: windows that contain an ambiguous base (a.amb, one bit per base) produce nothing
// (run_skip_ambiguous_windows, src/lib.rs:451-496); a separate instance so that the plain path
// carries no extra state.
// XW: a.w > FAST_MAX_W.  The van-Herk machinery computes the minima of sub-windows of W k-mers;
// they go through a small per-warp ring of rows in the L2 scratch, and the minimum of the real
// window is the minimum over the T+1 shifted sub-windows that cover it (the current one from
// registers, the others read back from the ring a group of four k-mers ahead of use).  Everything
// downstream (flags, position bytes, strand fix-up, emission) sees the real-window result.
template <int W, bool HC, bool LR, bool SYNC, bool AMB = false, bool XW = false>
__global__ void __launch_bounds__(FAST_NT, FAST_BPS) mz_fast_kernel(const KArgs a) {
    static_assert(W >= 1 && W <= (int)FAST_MAX_W, "W out of range");
    constexpr int B = (int)fast_b(W);    // van-Herk blocks per loop iteration
    constexpr int SB = B * W;            // k-mers per loop iteration (<= 32)
    constexpr int WQ = (SB + 3) / 4;     // record words per iteration
    constexpr uint32_t NT = FAST_NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint4* T = reinterpret_cast<uint4*>(smem_raw);
    uint32_t* misc = reinterpret_cast<uint32_t*>(T + 256 * FAST_TC);
    const uint32_t LCAP = a.list_cap;
    uint32_t* const list = misc + 8 + warp * LCAP;  // this warp's staging list
    const uint32_t Wr = XW ? a.w : (uint32_t)W;  // real window length
    const uint32_t NBmax = fast_nb(a.S, Wr);
    // flag words of this warp, two tiles in flight: fl0[buf][b*32 + lane]
    uint32_t* const fl0 = misc + 8 + FAST_WARPS * LCAP + (size_t)warp * 2 * NBmax * 32;
    // this warp's record rows in global scratch (two buffers): row r, lane t -> sc0[buf][r*32 + t]
    uint32_t* const sc0 = a.scratch + ((size_t)blockIdx.x * FAST_WARPS + warp) * (2 * a.scratch_words_per_block + a.r1_words_per_warp);
    uint32_t* const r1 = sc0 + 2 * a.scratch_words_per_block;  // XW: level-1 rows of this warp

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
    uint32_t one;
    asm volatile("mov.u32 %0, 1;" : "=r"(one));  // opaque constant 1 for imad()
    // syncmer offsets d = (window end) - (selected pos): closed {0, W-1}, open {(W-1)/2}
    const uint32_t so1 = a.mode == MODE_CLOSED ? 0u : (Wr - 1) / 2, so2 = a.mode == MODE_CLOSED ? Wr - 1 : (Wr - 1) / 2;

    __syncthreads();  // table + misc ready; from here on every warp works on its own
    const bool minim = a.mode == MODE_MINIMIZER;

    // Software pipeline over tiles: the look-back + emission of tile A runs after the main loop
    // of the next tile B, so A's predecessors have published their counts by then.
    uint32_t p_valid = 0, p_tile = 0, p_cnt = 0, p_NB = 0, p_inc = 0, cur = 0;
    for (;;) {  // persistent: one tile (32 threads x S windows) per iteration, per warp
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(a.ticket, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        const bool have = tile < a.num_tiles;
        uint32_t cnt = 0, NB = 0, inc = 0;
        if (have) {
        const Segment sg = make_segment_nt(a, tile, lane, 32u);
        uint32_t* const scr = sc0 + (size_t)cur * a.scratch_words_per_block + lane;
        uint32_t* const flp = fl0 + (size_t)cur * NBmax * 32 + lane;

        if (sg.nvalid) {
            uint32_t fw = misc[1], rc = misc[2];
            const uint32_t nelem = sg.nvalid + sg.has_prev + (Wr - 1);
            NB = (nelem + SB - 1) / SB;
            // first/last valid window-end element: e = jl + W - 1, jl in [has_prev, has_prev + nvalid)
            const uint32_t e_lo = sg.has_prev + (Wr - 1), e_hi = e_lo + sg.nvalid;

            // ---- thread-local view of the packed stream (32-bit word offsets) ----------------
            // Entering bases of block b start at local base k-1+bW; leaving bases at local base
            // bW-1 (base -1 = virtual 'A').  Both are addressed with one extra virtual word in
            // front (+32 bits) so that the bit position never goes negative.
            const uint64_t w0abs = sg.bit0 >> 5;
            const uint32_t* const wbase = a.seq + w0abs;
            const uint32_t sh0 = (uint32_t)sg.bit0 & 31u;
            const uint64_t remw = a.seq_nwords - 1 - w0abs;
            const uint32_t wlim = remw > 0x7ffffff0ull ? 0x7ffffff0u : (uint32_t)remw;
            // ---- prologue: consume k-1 bases, two per table step (leaving bases = virtual 'A') --
            {
                uint32_t pp = sh0, rem = k - 1;
                while (rem) {
                    const uint32_t wl = pp >> 5, sh = pp & 31u;
                    uint32_t x = __funnelshift_r(__ldg(wbase + min(wl, wlim)), __ldg(wbase + min(wl + 1, wlim)), sh);
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
            // may any (pre)fetch of this thread touch a word past the end of the buffer?
            const bool clampd = ((pin + 2u * (NB + 1) * SB) >> 5) + 2u > wlim;
            auto ldw = [&](uint32_t wl1) -> uint32_t {  // wl1 = word offset + 1 (virtual word 0)
                if (wl1 == 0) return 0u;
                uint32_t wl = wl1 - 1;
                if (clampd) wl = min(wl, wlim);
                return __ldg(wbase + wl);
            };
            uint32_t iw0 = ldw(pin >> 5), iw1 = ldw((pin >> 5) + 1), iw2 = SB > 16 ? ldw((pin >> 5) + 2) : 0u;
            uint32_t ow0 = ldw(pout >> 5), ow1 = ldw((pout >> 5) + 1), ow2 = SB > 16 ? ldw((pout >> 5) + 2) : 0u;
            ow0 &= ~(3u << (pout & 31u));  // element 0 leaves the virtual 'A'

            uint32_t RL[W], RR[W];
#pragma unroll
            for (int t = 0; t < W; t++) RL[t] = 0xffffffffu, RR[t] = 0u;
            uint32_t prev = 0xffffffffu, prevlow = 0x100u;
            uint32_t* sp = scr;
            // XW: window of Wr k-mers ending at e = min over the sub-windows (W k-mers) ending at
            // e, e - W, ..., e - (T-1) W and e - (Wr - W); consecutive ones overlap or abut
            const uint32_t Dmax = Wr - W, T = XW ? (Wr + W - 1) / W - 1 : 0u;
            const uint32_t rmask = XW ? fast_ring_rows(Wr) - 1u : 0u;
            uint32_t tL[4] = {0, 0, 0, 0}, tR[4] = {0, 0, 0, 0};
            // ambiguity: only threads whose stretch holds an ambiguous base do any work per block
            bool amb_here = false;
            uint32_t zrun = 0, pclean = 0;
            uint64_t ab0 = 0;
            if (AMB) {
                ab0 = (uint64_t)((int64_t)sg.pos_base + a.amb_bitbias);
                amb_here = amb_any(a, ab0, nelem + k - 1);
                if (amb_here) zrun = amb_clean_run(a, ab0, k - 1);
            }

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
                uint32_t preL = 0, preR = 0, bf = 0, lastR = 0;
                uint32_t accL[WQ], accR[WQ];
#pragma unroll
                for (int q = 0; q < WQ; q++) accL[q] = 0, accR[q] = 0;
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
                    const uint32_t word = N[(t >> 4) * 2 + ((t >> 1) & 1)];
                    const uint32_t idx = get_byte(word, (t & 15) >> 2);
                    const uint32_t addr = idx * (16u * FAST_TC) + tb;
                    uint32_t h0, h1 = 0;
                    if (HC) {
                        const uint4 e = lds128(addr);
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
                        const uint2 e = lds64(addr);
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
                    if (LR) {
                        accR[(o + t) >> 2] = put_byte(accR[(o + t) >> 2], mR0, (o + t) & 3);
                        if (two) {
                            accR[(o + t + 1) >> 2] = put_byte(accR[(o + t + 1) >> 2], mR1, (o + t + 1) & 3);
                        }
                        if (o + t == SB - 1) lastR = mR0;
                        if (o + t + 1 == SB - 1) lastR = mR1;
                    }
                    if (SYNC) {
                        const uint32_t d0 = pos0 - (res0 & 0xffffu);
                        if (d0 == so1 || d0 == so2) bf |= 1u << (o + t);
                        if (two) {
                            const uint32_t d1 = pos1 - (res1 & 0xffffu);
                            if (d1 == so1 || d1 == so2) bf |= 1u << (o + t + 1);
                        }
                    } else {
                        or_if_ne(bf, res0, prev, 1u << (o + t));
                        prev = res0;
                        if (two) {
                            or_if_ne(bf, res1, prev, 1u << (o + t + 1));
                            prev = res1;
                        }
                    }
                    accL[(o + t) >> 2] = put_byte(accL[(o + t) >> 2], res0, (o + t) & 3);
                    if (two) accL[(o + t + 1) >> 2] = put_byte(accL[(o + t + 1) >> 2], res1, (o + t + 1) & 3);
                }
                // suffix minima of this block (slot 0 is never needed)
#pragma unroll
                for (int q = W - 2; q >= 1; q--) {
                    RL[q] = min(RL[q], RL[q + 1]);
                    if (LR) RR[q] = max(RR[q], RR[q + 1]);
                }
                }  // van-Herk blocks of this iteration
                // leftmost != rightmost somewhere in this iteration?  Compare the position bytes,
                // four windows per LOP3 (positions inside a window differ by < 256)
                uint32_t tacc = 0;
                if (LR) {
#pragma unroll
                    for (int q = 0; q < WQ; q++) tacc |= accL[q] ^ accR[q];
                }
                if (LR && tacc != 0u) {
                    // cold: some window of this block has leftmost != rightmost.  Apply the strand
                    // rule to those windows and rebuild the block's flags from the position bytes.
                    bf = 0;
                    uint32_t pl = prevlow;
#pragma unroll
                    for (int t = 0; t < SB; t++) {
                        uint32_t cur = get_byte(accL[t >> 2], t & 3);
                        const uint32_t rgt = get_byte(accR[t >> 2], t & 3);
                        if (cur != rgt && eb + t >= Wr - 1u) {
                            if (!window_prefers_left(wbase, sh0, wlim, 2u * (eb + t - (Wr - 1u)), a.l)) {
                                cur = rgt;
                                accL[t >> 2] = put_byte(accL[t >> 2], rgt, t & 3);
                                if (t == SB - 1) prev = lastR ^ 0xffff0000u;
                            }
                        }
                        if (SYNC) {
                            const uint32_t d = (eb + t - cur) & 0xffu;
                            if (d == so1 || d == so2) bf |= 1u << t;
                        } else {
                            if (cur != pl) bf |= 1u << t;
                            pl = cur;
                        }
                    }
                }
                prevlow = get_byte(accL[(SB - 1) >> 2], (SB - 1) & 3);
                uint32_t clean = 0xffffffffu;
                if (AMB && amb_here) {
                    // window ending at k-mer eb+t = the l bases ending at local base eb+t+k-1.
                    // A clean window right after an ambiguous one is always emitted: the
                    // reference compares against SKIPPED there (src/intrinsics/dedup.rs:147-155).
                    const uint2 cm = amb_clean_mask(a, ab0 + eb + (k - 1), SB, a.l, zrun);
                    clean = cm.x;
                    zrun = cm.y;
                    if (!SYNC) bf |= ~((clean << 1) | pclean);
                    pclean = (clean >> (SB - 1)) & 1u;
                }
                // keep flags of valid windows only: bit t <-> window-end element eb + t
                // (only the first and last blocks of a segment can hold invalid windows)
                if (eb < e_lo + 1u || eb + SB > e_hi) {
                    const uint32_t lo = e_lo > eb ? min(e_lo - eb, (uint32_t)SB) : 0u;
                    const uint32_t hi = e_hi > eb ? min(e_hi - eb, (uint32_t)SB) : 0u;
                    const uint32_t mhi = hi >= 32u ? 0xffffffffu : ((1u << hi) - 1u);
                    const uint32_t mlo = lo >= 32u ? 0xffffffffu : ((1u << lo) - 1u);
                    if (!SYNC && sg.first_always && e_lo >= eb && e_lo < eb + SB) bf |= 1u << (e_lo - eb);
                    bf &= mhi & ~mlo;
                }
                if (AMB) bf &= clean;
                flp[b * 32] = bf;
#pragma unroll
                for (int q = 0; q < WQ; q++) sp[q * 32] = accL[q];
                sp += WQ * 32;
                cnt += __popc(bf);
            }
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
        }  // have

        // ---- ordered emission of the PREVIOUS tile, warp-autonomous (no block barriers) --------
        if (p_valid) {
        const uint32_t tile_e = p_tile, cnt_e = p_cnt, NB_e = p_NB, inc_e = p_inc;
        (void)NB_e;
        const uint32_t* const scr0 = pinned(sc0 + (size_t)(cur ^ 1u) * a.scratch_words_per_block);
        const uint32_t* const flr = fl0 + (size_t)(cur ^ 1u) * NBmax * 32 + lane;
        const uint32_t total = __shfl_sync(0xffffffffu, inc_e, 31);
        const uint32_t toff = inc_e - cnt_e;
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

        // Tile-uniform addressing (32-bit offsets from the tile's first thread).
        const bool canon_val = a.val_canonical != 0;
        const uint64_t j00 = a.wbeg + (uint64_t)tile_e * 32u * a.S;  // first window of the tile
        const uint32_t hp0 = (a.n_reads == 0 && j00 > 0 && minim) ? 1u : 0u;
        const uint64_t tbit0 = (uint64_t)((int64_t)(2 * (j00 - hp0)) + a.bitbias);  // lane 0's bit0
        const uint32_t* const twbase = pinned(a.seq + (tbit0 >> 5));
        const uint32_t tsh = (uint32_t)tbit0 & 31u;
        const uint64_t trem = a.seq_nwords - 1 - (tbit0 >> 5);
        const uint32_t twlim = trem > 0x7ffffff0ull ? 0x7ffffff0u : (uint32_t)trem;
        const bool tclamp = ((tsh + 2u * (32u * a.S + a.l + 64u)) >> 5) + 3u > twlim;
        const uint32_t pos00 = (uint32_t)j00;

        // Every lane walks its own flag words (coalesced across lanes) and stages its entries at
        // [toff, toff + cnt) of the tile's output range; list_cap entries per pass.
        uint32_t bq = 0, produced = 0;
        uint32_t f = (cnt_e != 0) ? flr[0] : 0u;
        for (uint32_t cbase = 0; cbase < total; cbase += LCAP) {
            while (produced < cnt_e && toff + produced < cbase + LCAP) {
                while (f == 0) {
                    bq++;
                    f = flr[bq * 32];
                }
                const uint32_t bit = (uint32_t)__ffs(f) - 1u;
                f &= f - 1u;
                list[toff + produced - cbase] = (lane << 16) | (bq << 5) | bit;  // owner | iteration | bit
                produced++;
            }
            __syncwarp();
            const uint32_t nent = min(LCAP, total - cbase);
            uint32_t* const opos = pinned(a.pos + (gbase + cbase));
            uint32_t* const osk = pinned(a.sk + (a.want_sk ? gbase + cbase : 0ull));
            unsigned long long* const oval = pinned(reinterpret_cast<unsigned long long*>(a.val) + (a.value_bits == 64 ? gbase + cbase : 0ull));
            if (a.n_reads == 0 && a.value_bits != 0) {
            // Single sequence with values: one software-pipelined pass over the staged entries (entry x is
            // handled by lane x & 31 in all stages, so nothing goes back through shared memory).
            //   stage A (entry x+64): descriptor -> load of the position byte from the L2 scratch
            //   stage B (entry x+32): position -> loads of the k-mer's words from the sequence
            //   stage C (entry x):    value, coalesced streaming stores of pos / sk / val
            // Both dependent load latencies are overlapped with the work of the previous entries.
            const bool want64 = a.value_bits == 64;
            uint32_t dA = 0, wvA = 0;                              // stage A -> B
            uint32_t relB = 0, skB = 0, w0 = 0, w1 = 0, w2 = 0;    // stage B -> C
            auto stageA = [&](uint32_t x) {
                dA = list[x];
                const uint32_t t2 = dA >> 16, b2 = (dA >> 5) & 0x7ffu, bit2 = dA & 31u;
                wvA = __ldcg(scr0 + (size_t)(b2 * WQ + (bit2 >> 2)) * 32 + t2);
            };
            auto stageB = [&]() {
                const uint32_t t2 = dA >> 16, b2 = (dA >> 5) & 0x7ffu, bit2 = dA & 31u, e = b2 * SB + bit2;
                const uint32_t lowb = (wvA >> (8u * (bit2 & 3u))) & 0xffu;
                const uint32_t jl = e - (Wr - 1);  // local window index of the owner
                const uint32_t local = minim ? jl + ((lowb - jl) & 0xffu) : jl;
                // owner's local base 0 = tile base + t2*S (- has_prev, which only differs for
                // the very first thread of the sequence)
                const uint32_t hp = (minim && !(j00 == 0 && t2 == 0)) ? 1u : 0u;
                relB = t2 * a.S - hp + hp0 + local;  // bases from the tile's bit0
                skB = pos00 + t2 * a.S + (jl - hp);
                if (want64) {
                    const uint32_t wl = (tsh + 2u * relB) >> 5;
                    if (!tclamp) {
                        const uint32_t* pw = twbase + wl;
                        w0 = __ldg(pw), w1 = __ldg(pw + 1), w2 = __ldg(pw + 2);
                    } else {
                        w0 = __ldg(twbase + min(wl, twlim)), w1 = __ldg(twbase + min(wl + 1, twlim));
                        w2 = __ldg(twbase + min(wl + 2, twlim));
                    }
                }
            };
            uint32_t x = lane;
            if (x < nent) {
                stageA(x);
                stageB();
            }
            if (x + 32 < nent) stageA(x + 32);
#pragma unroll 1
            for (; x < nent; x += 32) {
                const uint32_t rel = relB, skv = skB, c0 = w0, c1 = w1, c2 = w2;
                if (x + 32 < nent) stageB();
                if (x + 64 < nent) stageA(x + 64);
                __stcs(opos + x, pos00 - hp0 + rel);  // streaming store: keep the L2 for the scratch rows
                if (a.want_sk) __stcs(osk + x, skv);
                if (want64) {
                    const uint32_t sh = (tsh + 2u * rel) & 31u;
                    const uint32_t vlo = __funnelshift_r(c0, c1, sh), vhi = __funnelshift_r(c1, c2, sh);
                    uint64_t v = (uint64_t)vlo | ((uint64_t)vhi << 32);
                    const uint32_t len = a.val_len;
                    if (len < 32) v &= (1ull << (2 * len)) - 1ull;
                    if (canon_val) {
                        const uint64_t r = (swap_pairs64(__brevll(v)) ^ 0xAAAAAAAAAAAAAAAAull) >> (64 - 2 * len);
                        v = r < v ? r : v;
                    }
                    __stcs(oval + x, (unsigned long long)v);
                } else if (a.value_bits == 128) {
                    uint64_t lo, hi;
                    kmer_value_u128(a, tbit0 + 2ull * rel, a.val_len, canon_val, lo, hi);
                    reinterpret_cast<ulonglong2*>(a.val)[gbase + cbase + x] = make_ulonglong2(lo, hi);
                }
            }
            } else {
            // positions only (nothing to overlap: measured 2-3 % faster this way), batch mode: two sweeps
            // sweep 1: fetch each entry's position byte from the L2 scratch (independent loads,
            // unrolled so several are in flight) and fold it into the staged descriptor
#pragma unroll 2
            for (uint32_t x = lane; x < nent; x += 32) {
                const uint32_t dsc = list[x];
                const uint32_t t2 = dsc >> 16, b2 = (dsc >> 5) & 0x7ffu, bit2 = dsc & 31u, e = b2 * SB + bit2;
                const uint32_t wv = __ldcg(scr0 + (size_t)(b2 * WQ + (bit2 >> 2)) * 32 + t2);
                const uint32_t lowb = (wv >> (8u * (bit2 & 3u))) & 0xffu;
                const uint32_t jl = e - (Wr - 1);  // local window index of the owner
                list[x] = (t2 << 27) | (((lowb - jl) & 0xffu) << 16) | jl;
            }
            __syncwarp();
            // sweep 2: positions, super-k-mer starts and k-mer values
#pragma unroll 1
            for (uint32_t x = lane; x < nent; x += 32) {
                const uint32_t dsc = list[x];
                const uint32_t t2 = dsc >> 27, jl = dsc & 0xffffu;
                const uint32_t local = minim ? jl + ((dsc >> 16) & 0xffu) : jl;
                if (a.n_reads == 0) {
                    const uint32_t hp = (minim && !(j00 == 0 && t2 == 0)) ? 1u : 0u;
                    __stcs(opos + x, pos00 + t2 * a.S - hp + local);
                    if (a.want_sk) __stcs(osk + x, pos00 + t2 * a.S + (jl - hp));
                    continue;
                }
                const Segment og = make_segment_nt(a, tile_e, t2, 32u);
                const unsigned long long oi = gbase + cbase + x;
                a.pos[oi] = (uint32_t)og.pos_base + local;
                if (a.want_sk) a.sk[oi] = (uint32_t)og.win_base + (jl - og.has_prev);
                if (a.value_bits == 64) {
                    a.val[oi] = kmer_value_u64(a, og.bit0 + 2ull * local, a.val_len, canon_val);
                } else if (a.value_bits == 128) {
                    uint64_t lo, hi;
                    kmer_value_u128(a, og.bit0 + 2ull * local, a.val_len, canon_val, lo, hi);
                    reinterpret_cast<ulonglong2*>(a.val)[oi] = make_ulonglong2(lo, hi);
                }
            }
            }
            __syncwarp();
        }
        }  // !ovf && total
        }  // p_valid
        if (!have) break;
        p_valid = 1, p_tile = tile, p_cnt = cnt, p_NB = NB, p_inc = inc;
        cur ^= 1u;
    }
}

// ---- host side ---------------------------------------------------------------------------
struct FastPlan {
    uint32_t S = 0, num_tiles = 0, grid = 0, list_cap = 0;
    size_t scratch_words_per_block = 0;  // position-byte rows, per warp and buffer
    size_t r1_words = 0;                 // XW: level-1 rows, per warp
};

// Geometry for the fast kernel; returns false when (k, w, ...) is outside its domain.
inline bool plan_fast(int sm_count, const mz_params& p, uint64_t nwin, FastPlan* pl, bool allow_xw = false) {
    const bool xw = p.w > FAST_MAX_W;
    if (xw && (!allow_xw || p.w > FAST_XW_MAX_W || getenv("MZ_NO_XW"))) return false;
    const char* env_s = getenv("MZ_FAST_S");
    const char* env_bps = getenv("MZ_FAST_BPS");
    const uint32_t bps = env_bps ? (uint32_t)atoi(env_bps) : FAST_BPS;  // resident blocks per SM
    const uint64_t slots = (uint64_t)sm_count * bps * FAST_WARPS;  // resident warps
    uint32_t s;
    if (env_s) {
        s = (uint32_t)atoi(env_s);
    } else {
        // long segments amortise the (k+w-2)-base warm-up.  Few waves (small inputs, shards of a
        // multi-GPU run): equal tiles in a whole number of waves, m = 1 .. 8 tiles per warp.
        const uint32_t smax = xw ? 310u + p.w : 310u;
        const uint64_t per_wave = slots * 32;
        const uint64_t m = (nwin + per_wave * smax - 1) / (per_wave * smax);
        const uint64_t want = m <= 8 ? (nwin + per_wave * m - 1) / (per_wave * m) : smax;
        s = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(want, 64), smax);
        // a thread computes S + w k-mers in whole loop iterations of SB k-mers: pick S so that
        // the last iteration is full (S = NB*SB - w; 288 -> 304 for w = 19 was worth 3 %)
        const uint32_t sb = fast_sb(fast_wt(p.w));
        uint32_t nb = std::max<uint32_t>(1, (s + p.w + (m <= 8 ? sb - 1 : 0)) / sb);  // few waves: round up
        while (nb * sb < p.w + 16) nb++;
        s = nb * sb - p.w;
        // dense outputs: shorten the segments until a tile's expected entries fit one staging
        // pass (a second pass costs more than the longer halo), but keep S >= 160
        const double dens = p.mode == MZ_MODE_MINIMIZER ? 2.0 / (p.w + 1.0)
                          : p.mode == MZ_MODE_CLOSED_SYNCMER ? (p.w == 1 ? 1.0 : 2.0 / p.w) : 1.0 / p.w;
        while (nb > 1 && (nb - 1) * sb >= p.w + 160 && 32.0 * s * dens * 1.15 > fast_list_cap(s, p)) {
            nb--;
            s = nb * sb - p.w;
        }
    }
    s = std::max<uint32_t>(16, s);
    // flag words and staging lists live in shared memory: stay within FAST_SMEM_LIMIT
    while (s > 16 + fast_sb(fast_wt(p.w)) && fast_smem(s, p.w, fast_list_cap(s, p)) > FAST_SMEM_LIMIT) s -= fast_sb(fast_wt(p.w));
    if ((uint64_t)s + p.w + 2 >= 65535 || fast_nb(s, p.w) >= 2048) return false;  // descriptor: 11-bit iteration
    const uint64_t Tt = (uint64_t)32 * s;
    const uint64_t tiles = (nwin + Tt - 1) / Tt;
    if (tiles == 0 || tiles > 0x7fffffffull) return false;
    pl->S = s;
    pl->list_cap = fast_list_cap(s, p);
    pl->num_tiles = (uint32_t)tiles;
    // one block per SM as soon as there are that many tiles (warps take tiles from the ticket
    // counter, so a small launch spreads over all SMs instead of filling a few of them)
    pl->grid = (uint32_t)std::min<uint64_t>(tiles, (uint64_t)sm_count * bps);
    pl->scratch_words_per_block = fast_scratch_words(s, p.w);
    pl->r1_words = fast_r1_words(s, p.w, p.strand_tiebreak != 0);
    return true;
}

template <int W, bool HC, bool LR, bool SYNC, bool AMB = false, bool XW = false>
inline int launch_fast_inst(uint32_t grid, const KArgs& a, cudaStream_t st) {
    auto kern = mz_fast_kernel<W, HC, LR, SYNC, AMB, XW>;
    const size_t smem = fast_smem(a.S, XW ? a.w : (uint32_t)W, a.list_cap);
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
