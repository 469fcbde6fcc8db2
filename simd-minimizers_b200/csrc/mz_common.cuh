// mz_common.cuh -- shared device helpers for the minimizer kernels (sm_100a).
//
// Semantics follow rust-seq/simd-minimizers v3.0.0 (citations relative to the reference tree):
//   key = hash >> 16, leftmost / rightmost argmin   src/sliding_min.rs:102-105,126-129,190-195
//   strand rule 2*#TG > l                           src/canonical.rs:19-29
//   dedup + super-k-mer index                       src/collect.rs:39-76
//   syncmer predicates                              src/syncmers.rs:33-37
//   k-mer values                                    src/lib.rs:598-629
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mz {

enum : uint32_t { MODE_MINIMIZER = 0, MODE_CLOSED = 1, MODE_OPEN = 2 };

// Kernel arguments (POD, passed by value).
struct KArgs {
    const uint32_t* seq;  // packed 2-bit bases as 32-bit words (16 bases per word)
    uint64_t seq_nwords;  // words readable at seq (loads are clamped to this)
    int64_t bitbias;      // bit position of global base g inside seq = 2*g + bitbias
    uint64_t nwin;        // windows of the whole sequence: n - l + 1
    uint64_t wbeg, wend;  // window range produced by this launch
    uint32_t k, w, l;
    uint32_t S;           // windows per thread
    uint32_t mode, want_sk, value_bits, val_len;
    uint32_t val_canonical;  // values are min(kmer, revcomp) (Builder CANONICAL, src/lib.rs:602)
    uint32_t f[4], c[4], rot;
    uint32_t* pos;
    uint32_t* sk;
    uint64_t* val;
    uint64_t cap;
    unsigned long long* tile_state;  // decoupled look-back descriptors, zeroed before launch
    uint32_t* ticket;                // dynamic tile id counter, zeroed before launch
    unsigned long long* count_out;   // total entries produced by this launch
    // the same two results written straight into mapped pinned host memory (no D2H copy node
    // after the kernel: one stream operation less per launch); zeroed by the host before launch
    unsigned long long* h_count;
    uint32_t* h_overflow;
    uint32_t* overflow;              // set to 1 when cap was too small
    uint32_t num_tiles;
    uint32_t* scratch;                 // fast kernel: per-warp record rows (L2-resident); generic: global ring
    uint64_t scratch_words_per_block;  // fast kernel: words per WARP and buffer; generic: per block
    uint32_t list_cap;                 // (generic kernel only)
    uint64_t r1_words_per_warp;        // fast kernel, w > 32: level-1 result rows per warp (else 0)
    // fast kernel: per-lane emission queues in shared memory (rows of 32 entries, one per lane)
    uint32_t q_rows;                   // rows per warp and buffer (q_trig + one loop iteration of guard rows)
    uint32_t q_bufs;                   // queue buffers per warp: 2 = emission deferred by one tile, 1 = right away
    uint32_t q_trig;                   // a lane holding more rows than this at the end of an iteration spills the warp's queues
    uint32_t nb;                       // loop iterations per tile (the same for every lane)
    uint32_t lead;                     // elements in front of the first valid window end
    uint32_t one;                      // always 1 (see mz_fast.cuh: keeps additions on the FMA pipe)
    // batch mode (thread per read); reads == 0 -> single sequence
    uint64_t n_reads;
    const uint64_t* read_start_bp;   // may be null -> fixed stride
    const uint32_t* read_len_bp;
    uint64_t stride_bits;            // fixed stride between reads, in bits
    uint32_t fixed_len_bp;
    uint64_t* out_offsets;           // CSR offsets, n_reads + 1
    // long reads: a thread handles one PIECE (S windows) of a read; n_reads then counts pieces
    const uint32_t* piece_read;      // read index of every piece (null: piece == read)
    const uint32_t* piece_win0;      // first window of the piece inside its read
    // ambiguous bases (PackedNSeq, src/lib.rs:451-496): one bit per base, null = none.
    // Windows containing an ambiguous base produce nothing (single-sequence mode only).
    const uint32_t* amb;             // bit of global base g = g + amb_bitbias
    uint64_t amb_nwords;
    int64_t amb_bitbias;
};

__device__ __forceinline__ uint32_t rotl32(uint32_t x, uint32_t r) {
    return __funnelshift_l(x, x, r);
}
__device__ __forceinline__ uint32_t rotr32(uint32_t x, uint32_t r) {
    return __funnelshift_r(x, x, r);
}

__device__ __forceinline__ uint32_t ld_word(const KArgs& a, uint64_t widx) {
    // clamp: bases past the end only ever feed windows that are discarded
    widx = widx < a.seq_nwords ? widx : a.seq_nwords - 1;
    return __ldg(a.seq + widx);
}

// 32 bits of the packed stream starting at absolute bit position `bit`
__device__ __forceinline__ uint32_t ld_bits32(const KArgs& a, uint64_t bit) {
    uint64_t wi = bit >> 5;
    uint32_t sh = (uint32_t)bit & 31u;
    uint32_t lo = ld_word(a, wi), hi = ld_word(a, wi + 1);
    return __funnelshift_r(lo, hi, sh);
}

// ---- ambiguity mask (one bit per base) -------------------------------------------------------
// (bits in front of / behind the mask read as 0 = unambiguous)
__device__ __forceinline__ uint32_t amb_word(const KArgs& a, int64_t wi) {
    return (wi >= 0 && (uint64_t)wi < a.amb_nwords) ? __ldg(a.amb + wi) : 0u;
}
// 32 mask bits starting at bit index `bit` (may be negative)
__device__ __forceinline__ uint32_t amb_bits32(const KArgs& a, int64_t bit) {
    const int64_t wi = bit >> 5;
    return __funnelshift_r(amb_word(a, wi), amb_word(a, wi + 1), (uint32_t)bit & 31u);
}
// any ambiguous base among the n bases starting at mask bit `bit`?
static __device__ __noinline__ bool amb_any(const KArgs& a, int64_t bit, uint32_t n) {
    for (; n >= 32; n -= 32, bit += 32)
        if (amb_bits32(a, bit)) return true;
    return n != 0 && (amb_bits32(a, bit) & ((1u << n) - 1u)) != 0;
}
// number of unambiguous bases at the end of the n bases starting at mask bit `bit`
static __device__ __noinline__ uint32_t amb_clean_run(const KArgs& a, int64_t bit, uint32_t n) {
    uint32_t run = 0;
    for (uint32_t i = 0; i < n; i += 32) {
        const uint32_t m = amb_bits32(a, bit + i), take = n - i < 32u ? n - i : 32u;
        for (uint32_t t = 0; t < take; t++) run = ((m >> t) & 1u) ? 0u : run + 1u;
    }
    return run;
}
// One loop iteration of the fast kernel: bit t of .x <=> the l bases ending at mask bit
// `bit + t` are all unambiguous (t < n <= 32); `run` = clean run ending just before `bit`,
// .y = the run after the n bases.
static __device__ __noinline__ uint2 amb_clean_mask(const KArgs& a, int64_t bit, uint32_t n, uint32_t l,
                                                    uint32_t run) {
    const uint32_t m = amb_bits32(a, bit);
    uint32_t clean = 0;
    for (uint32_t t = 0; t < n; t++) {
        run = ((m >> t) & 1u) ? 0u : run + 1u;
        if (run >= l) clean |= 1u << t;
    }
    return make_uint2(clean, run);
}

// Sequential reader of 2-bit bases from a bit position.
struct BaseReader {
    uint64_t widx;
    uint32_t cur;
    int left;
    __device__ __forceinline__ void init(const KArgs& a, uint64_t bit) {
        widx = bit >> 5;
        uint32_t sh = (uint32_t)bit & 31u;
        cur = ld_word(a, widx) >> sh;
        left = (int)((32u - sh) >> 1);
    }
    __device__ __forceinline__ uint32_t next(const KArgs& a) {
        if (left == 0) {
            widx++;
            cur = ld_word(a, widx);
            left = 16;
        }
        uint32_t b = cur & 3u;
        cur >>= 2;
        left--;
        return b;
    }
};

// number of T/G bases (code & 2) among `len` bases starting at bit position `bit`
__device__ __forceinline__ uint32_t tg_count(const KArgs& a, uint64_t bit, uint32_t len) {
    uint32_t cnt = 0;
    uint32_t bits = 2 * len;
    while (bits >= 32) {
        cnt += __popc(ld_bits32(a, bit) & 0xAAAAAAAAu);
        bit += 32;
        bits -= 32;
    }
    if (bits) cnt += __popc(ld_bits32(a, bit) & 0xAAAAAAAAu & ((1u << bits) - 1u));
    return cnt;
}

__device__ __forceinline__ uint64_t swap_pairs64(uint64_t r) {
    return ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
}

// 2-bit packed k-mer of `len` <= 32 bases at bit position `bit`; canonical: min(kmer, revcomp)
__device__ __forceinline__ uint64_t kmer_value_u64(const KArgs& a, uint64_t bit, uint32_t len,
                                                   bool canonical) {
    uint64_t wi = bit >> 5;
    uint32_t sh = (uint32_t)bit & 31u;
    uint32_t w0 = ld_word(a, wi), w1 = ld_word(a, wi + 1), w2 = ld_word(a, wi + 2);
    uint64_t v = (uint64_t)__funnelshift_r(w0, w1, sh) | ((uint64_t)__funnelshift_r(w1, w2, sh) << 32);
    if (len < 32) v &= (1ull << (2 * len)) - 1ull;
    if (!canonical) return v;
    uint64_t r = swap_pairs64(__brevll(v)) ^ 0xAAAAAAAAAAAAAAAAull;
    r >>= (64 - 2 * len);
    return r < v ? r : v;
}

// len <= 64; result as (lo, hi)
__device__ __forceinline__ void kmer_value_u128(const KArgs& a, uint64_t bit, uint32_t len,
                                                bool canonical, uint64_t& lo, uint64_t& hi) {
    uint64_t wi = bit >> 5;
    uint32_t sh = (uint32_t)bit & 31u;
    uint32_t w[5];
#pragma unroll
    for (int i = 0; i < 5; i++) w[i] = ld_word(a, wi + i);
    uint64_t vlo = (uint64_t)__funnelshift_r(w[0], w[1], sh) | ((uint64_t)__funnelshift_r(w[1], w[2], sh) << 32);
    uint64_t vhi = (uint64_t)__funnelshift_r(w[2], w[3], sh) | ((uint64_t)__funnelshift_r(w[3], w[4], sh) << 32);
    if (len <= 32) {
        vhi = 0;
        if (len < 32) vlo &= (1ull << (2 * len)) - 1ull;
    } else if (len < 64) {
        vhi &= (1ull << (2 * (len - 32))) - 1ull;
    }
    lo = vlo;
    hi = vhi;
    if (!canonical) return;
    // reverse the 128-bit value 2 bits at a time, complement, shift down by 128 - 2*len
    uint64_t rhi = swap_pairs64(__brevll(vlo)) ^ 0xAAAAAAAAAAAAAAAAull;
    uint64_t rlo = swap_pairs64(__brevll(vhi)) ^ 0xAAAAAAAAAAAAAAAAull;
    uint32_t s = 128 - 2 * len;  // 0..126, even
    uint64_t olo, ohi;
    if (s == 0) {
        olo = rlo, ohi = rhi;
    } else if (s < 64) {
        olo = (rlo >> s) | (rhi << (64 - s));
        ohi = rhi >> s;
    } else {
        olo = rhi >> (s - 64);  // s == 64 -> shift 0
        ohi = 0;
    }
    bool rc_smaller = (ohi < vhi) || (ohi == vhi && olo < vlo);
    if (rc_smaller) lo = olo, hi = ohi;
}

// ---------------------------------------------------------------------------------------------
// Decoupled look-back over tile descriptors: state = (value << 2) | status,
// status 0 = not ready, 1 = tile aggregate, 2 = inclusive prefix.  Called by warp 0 of a block.
// Returns the exclusive prefix of `tile` (sum of totals of all earlier tiles).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_state(unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long lookback_warp0(unsigned long long* state,
                                                              uint32_t tile,
                                                              unsigned long long total) {
    const uint32_t lane = threadIdx.x & 31u;
    if (tile == 0) {
        if (lane == 0) st_state(state, (total << 2) | 2ull);
        return 0;
    }
    if (lane == 0) st_state(state + tile, (total << 2) | 1ull);
    unsigned long long excl = 0;
    int64_t base = (int64_t)tile - 1;
    while (true) {
        int64_t idx = base - (int64_t)lane;
        unsigned long long v = 2ull;  // virtual tile before 0: inclusive prefix 0
        if (idx >= 0) {
            v = ld_state(state + idx);
            while ((v & 3ull) == 0ull) {
                __nanosleep(64);
                v = ld_state(state + idx);
            }
        }
        uint32_t is_prefix = __ballot_sync(0xffffffffu, (v & 3ull) == 2ull);
        uint32_t first = is_prefix ? (uint32_t)__ffs(is_prefix) - 1u : 32u;
        unsigned long long contrib = (lane <= first) ? (v >> 2) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        excl += contrib;
        if (is_prefix) break;
        base -= 32;
    }
    if (lane == 0) st_state(state + tile, ((excl + total) << 2) | 2ull);
    return excl;
}

// Variant for pipelined tiles: the tile's aggregate (status 1) has ALREADY been published; walk
// back to find the exclusive prefix, then publish the inclusive prefix.  Called by a whole warp.
__device__ __forceinline__ unsigned long long lookback_excl(unsigned long long* state, uint32_t tile,
                                                            unsigned long long total) {
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long excl = 0;
    if (tile != 0) {
        int64_t base = (int64_t)tile - 1;
        // First step: the 32 nearest predecessors (in steady state one of them already holds an
        // inclusive prefix).  If none does -- the first wave of a launch, where every tile's
        // prefix depends on all tiles before it -- continue 256 descriptors at a time, eight
        // independent loads per lane, so the chain is tile/256 L2 round trips instead of tile/32.
        constexpr int K = 8;
        bool found = false;
        {
            int64_t idx = base - (int64_t)lane;
            unsigned long long v = 2ull;  // virtual tile before 0: inclusive prefix 0
            if (idx >= 0) {
                v = ld_state(state + idx);
                while ((v & 3ull) == 0ull) {
                    __nanosleep(64);
                    v = ld_state(state + idx);
                }
            }
            const uint32_t is_prefix = __ballot_sync(0xffffffffu, (v & 3ull) == 2ull);
            const uint32_t first = is_prefix ? (uint32_t)__ffs(is_prefix) - 1u : 32u;
            unsigned long long contrib = (lane <= first) ? (v >> 2) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
            excl += contrib;
            found = is_prefix != 0u;
            base -= 32;
        }
        while (!found) {
            unsigned long long v[K];
#pragma unroll
            for (int j = 0; j < K; j++) {
                const int64_t idx = base - (int64_t)lane - 32 * j;
                v[j] = idx >= 0 ? ld_state(state + idx) : 2ull;
            }
#pragma unroll
            for (int j = 0; j < K; j++) {
                const int64_t idx = base - (int64_t)lane - 32 * j;
                while ((v[j] & 3ull) == 0ull) {
                    __nanosleep(64);
                    v[j] = ld_state(state + idx);
                }
            }
#pragma unroll
            for (int j = 0; j < K; j++) {
                if (!found) {
                    const uint32_t is_prefix = __ballot_sync(0xffffffffu, (v[j] & 3ull) == 2ull);
                    const uint32_t first = is_prefix ? (uint32_t)__ffs(is_prefix) - 1u : 32u;
                    unsigned long long contrib = (lane <= first) ? (v[j] >> 2) : 0ull;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                    excl += contrib;
                    found = is_prefix != 0u;
                }
            }
            base -= 32 * K;
        }
    }
    if (lane == 0) st_state(state + tile, ((excl + total) << 2) | 2ull);
    return excl;
}

// Block-wide exclusive scan of one uint32 per thread; returns exclusive prefix, sets total.
// `scratch` needs blockDim.x/32 + 1 words of shared memory.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* scratch,
                                                         uint32_t& total) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += t;
    }
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t x = lane < nwarps ? scratch[lane] : 0u;
        uint32_t xi = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, xi, o);
            if (lane >= (uint32_t)o) xi += t;
        }
        if (lane < nwarps) scratch[lane] = xi - x;
        if (lane == 31) scratch[32] = xi;
    }
    __syncthreads();
    total = scratch[32];
    return scratch[warp] + inc - v;
}

}  // namespace mz
