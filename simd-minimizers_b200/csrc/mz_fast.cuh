// mz_fast.cuh -- W-specialised, register-resident minimizer / syncmer kernel for sm_100a.
//
// One thread walks S consecutive windows.  Per van-Herk block of W k-mers (fully unrolled):
//   * 2W bits of the entering- and leaving-base streams are fetched and re-aligned with funnel
//     shifts, then interleaved so that every byte holds (in,in,out,out) of two consecutive bases;
//   * one LDS.128 from a 256-entry table returns the rolling-hash deltas of BOTH bases for the
//     forward and the reverse-complement hash (ntHash/mulHash are GF(2)-linear in the tables),
//     so a k-mer hash costs SHF+LOP3 per strand;
//   * (hash & 0xffff0000) | pos goes through a prefix-min / suffix-min pair whose W-entry suffix
//     array lives in registers (static indexing); the rightmost minimum uses max on the
//     complemented key, exactly the reference's packing (src/sliding_min.rs:190-195,336-341);
//   * the strand rule (src/canonical.rs) is evaluated only when leftmost != rightmost;
//   * the low byte of the selected position and a flag bit per window are recorded in shared
//     memory; mz_emit.cuh turns them into ordered, coalesced output.
#pragma once
#include "../../include/mz_b200.h"
#include <algorithm>
#include <cstdlib>
#include "mz_emit.cuh"

namespace mz {

constexpr uint32_t FAST_MAX_W = 32;

// Shared-memory layout (NT = blockDim.x, NB = van-Herk blocks per thread, WQ = ceil(W/4)):
//   uint4 T[256]; uint32 misc[8]; emit staging; uint32 flagw[NB][NT]; uint32 recw[NB*WQ][NT]
__host__ __device__ constexpr uint32_t fast_wq(uint32_t W) { return (W + 3) / 4; }
__host__ __device__ inline uint32_t fast_nb(uint32_t S, uint32_t W) {
    return (S + 1 + (W - 1) + W - 1) / W;  // elements = S + has_prev + W - 1
}
inline size_t fast_smem(uint32_t NT, uint32_t S, uint32_t W) {
    return 256 * 16 + 32 + EMIT_SMEM_BYTES + (size_t)fast_nb(S, W) * (1 + fast_wq(W)) * 4 * NT;
}

// 64 bits of the packed stream starting at bit position `bit`
__device__ __forceinline__ void ld_bits64(const KArgs& a, uint64_t bit, uint32_t& lo, uint32_t& hi,
                                          bool need_hi) {
    uint64_t wi = bit >> 5;
    uint32_t sh = (uint32_t)bit & 31u;
    uint32_t w0 = ld_word(a, wi), w1 = ld_word(a, wi + 1);
    lo = __funnelshift_r(w0, w1, sh);
    hi = 0;
    if (need_hi) {
        uint32_t w2 = ld_word(a, wi + 2);
        hi = __funnelshift_r(w1, w2, sh);
    }
}

template <int W, bool HC, bool LR, bool SYNC>
__global__ void __launch_bounds__(256) mz_fast_kernel(const KArgs a) {
    static_assert(W >= 1 && W <= (int)FAST_MAX_W, "W out of range");
    constexpr int WQ = (W + 3) / 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t NT = blockDim.x, tid = threadIdx.x;
    uint4* T = reinterpret_cast<uint4*>(smem_raw);
    uint32_t* misc = reinterpret_cast<uint32_t*>(T + 256);
    unsigned char* after = reinterpret_cast<unsigned char*>(misc + 8);
    const EmitSmem es = EmitSmem::carve(after);
    uint32_t* flagw = reinterpret_cast<uint32_t*>(after + EMIT_SMEM_BYTES);
    const uint32_t NBmax = fast_nb(a.S, W);
    uint32_t* recw = flagw + (size_t)NBmax * NT;

    const uint32_t k = a.k, R = a.rot & 31u, R2 = (2u * R) & 31u;
    // ---- table: index byte = in0 | in1<<2 | out0<<4 | out1<<6 (two consecutive bases) --------
    for (uint32_t idx = tid; idx < 256; idx += NT) {
        const uint32_t in0 = idx & 3u, in1 = (idx >> 2) & 3u, out0 = (idx >> 4) & 3u, out1 = idx >> 6;
        const uint32_t rk = (R * k) & 31u, rk1 = (R * (k - 1)) & 31u;
        const uint32_t f0 = a.f[in0] ^ rotl32(a.f[out0], rk), f1 = a.f[in1] ^ rotl32(a.f[out1], rk);
        const uint32_t c0 = rotl32(a.c[in0], rk1) ^ rotr32(a.c[out0], R);
        const uint32_t c1 = rotl32(a.c[in1], rk1) ^ rotr32(a.c[out1], R);
        T[idx] = make_uint4(f0, rotl32(f0, R) ^ f1, c0, rotr32(c0, R) ^ c1);
    }
    if (tid == 0) {
        misc[0] = atomicAdd(a.ticket, 1u);
        uint32_t fa = 0, ca = 0;  // hash state of the virtual all-'A' k-mer before every segment
        for (uint32_t j = 0; j < k; j++) {
            fa ^= rotl32(a.f[0], (R * j) & 31u);
            ca ^= rotl32(a.c[0], (R * j) & 31u);
        }
        misc[1] = fa;
        misc[2] = ca;
    }
    __syncthreads();
    const uint32_t tile = misc[0];
    const Segment sg = make_segment(a, tile, tid);

    uint32_t cnt = 0, NB = 0;
    if (sg.nvalid) {
        uint32_t fw = misc[1], rc = misc[2];
        // ---- prologue: k-1 bases, one at a time (leaving base = virtual 'A') -----------------
        {
            BaseReader in;
            in.init(a, sg.bit0);
            for (uint32_t u = 0; u + 1 < k; u++) {
                const uint4 e = T[in.next(a)];
                fw = rotl32(fw, R) ^ e.x;
                if (HC) rc = rotr32(rc, R) ^ e.z;
            }
        }
        const uint32_t nelem = sg.nvalid + sg.has_prev + (W - 1);
        NB = (nelem + W - 1) / W;
        uint32_t RL[W], RR[W];
#pragma unroll
        for (int t = 0; t < W; t++) RL[t] = 0xffffffffu, RR[t] = 0u;
        uint32_t prev = 0xffffffffu;
        // syncmer offsets d = (window end) - (selected pos): closed {0, W-1}, open {(W-1)/2}
        const uint32_t so1 = a.mode == MODE_CLOSED ? 0u : (W - 1) / 2, so2 = a.mode == MODE_CLOSED ? W - 1 : (W - 1) / 2;
        // first/last valid window-end element: e = jl + W - 1, jl in [has_prev, has_prev + nvalid)
        const uint32_t e_lo = sg.has_prev + (W - 1), e_hi = e_lo + sg.nvalid;
        const uint64_t bit_in0 = sg.bit0 + 2ull * (k - 1);

        for (uint32_t b = 0; b < NB; b++) {
            const uint32_t eb = b * W;
            // entering bases: local bases k-1+eb .. ; leaving bases: local bases eb-1 ..
            uint32_t x0, x1, y0, y1;
            ld_bits64(a, bit_in0 + 2ull * eb, x0, x1, W > 16);
            if (b == 0) {
                ld_bits64(a, sg.bit0, y0, y1, W > 16);
                y1 = __funnelshift_l(y0, y1, 2);  // shift the stream up by one base:
                y0 <<= 2;                         // slot 0 leaves the virtual 'A'
            } else {
                ld_bits64(a, sg.bit0 + 2ull * (eb - 1), y0, y1, W > 16);
            }
            uint32_t N[4];
            N[0] = (x0 & 0x0F0F0F0Fu) | ((y0 << 4) & 0xF0F0F0F0u);
            N[1] = ((x0 >> 4) & 0x0F0F0F0Fu) | (y0 & 0xF0F0F0F0u);
            if (W > 16) {
                N[2] = (x1 & 0x0F0F0F0Fu) | ((y1 << 4) & 0xF0F0F0F0u);
                N[3] = ((x1 >> 4) & 0x0F0F0F0Fu) | (y1 & 0xF0F0F0F0u);
            }
            uint32_t preL = 0, preR = 0, bf = 0, acc = 0;
            uint32_t hpair = 0;
#pragma unroll
            for (int t = 0; t < W; t++) {
                uint32_t h;
                if ((t & 1) == 0) {
                    const uint32_t word = N[(t >> 4) * 2 + ((t >> 1) & 1)];
                    const int byte = (t & 15) >> 2;
                    const uint32_t idx = byte == 0 ? (word & 0xffu)
                                       : byte == 3 ? (word >> 24)
                                                   : __byte_perm(word, 0, 0x4440 + byte);
                    if (HC) {
                        const uint4 e = T[idx];
                        const uint32_t fA = rotl32(fw, R) ^ e.x, rA = rotr32(rc, R) ^ e.z;
                        h = fA + rA;
                        if (t + 1 < W) {
                            fw = rotl32(fw, R2) ^ e.y;
                            rc = rotr32(rc, R2) ^ e.w;
                            hpair = fw + rc;
                        } else {
                            fw = fA;
                            rc = rA;
                        }
                    } else {
                        const uint2 e = *reinterpret_cast<const uint2*>(&T[idx]);
                        const uint32_t fA = rotl32(fw, R) ^ e.x;
                        h = fA;
                        if (t + 1 < W) {
                            fw = rotl32(fw, R2) ^ e.y;
                            hpair = fw;
                        } else {
                            fw = fA;
                        }
                    }
                } else {
                    h = hpair;
                }
                const uint32_t pos = eb + t;
                const uint32_t le = (h & 0xffff0000u) | pos;
                preL = t == 0 ? le : min(preL, le);
                uint32_t res = t < W - 1 ? min(preL, RL[t < W - 1 ? t + 1 : 0]) : preL;
                RL[t] = le;
                if (LR) {
                    const uint32_t re = le ^ 0xffff0000u;
                    preR = t == 0 ? re : max(preR, re);
                    const uint32_t mR = t < W - 1 ? max(preR, RR[t < W - 1 ? t + 1 : 0]) : preR;
                    RR[t] = re;
                    if (((res ^ mR) & 0xffffu) != 0u && pos >= (uint32_t)(W - 1)) {
                        // leftmost != rightmost: strand rule on the window's l bases
                        const uint32_t tg = tg_count(a, sg.bit0 + 2ull * (pos - (W - 1)), a.l);
                        if (!(2u * tg > a.l)) res = mR ^ 0xffff0000u;
                    }
                }
                bool flag;
                if (SYNC) {
                    const uint32_t d = pos - (res & 0xffffu);
                    flag = d == so1 || d == so2;
                } else {
                    flag = res != prev;
                    prev = res;
                }
                if (flag) bf |= 1u << t;
                acc = __byte_perm(acc, res, (t & 3) == 0 ? 0x3214 : (t & 3) == 1 ? 0x3240 : (t & 3) == 2 ? 0x3410 : 0x4210);
                if ((t & 3) == 3 || t == W - 1) recw[((size_t)b * WQ + (t >> 2)) * NT + tid] = acc;
            }
            // suffix minima of this block (slot 0 is never needed)
#pragma unroll
            for (int q = W - 2; q >= 1; q--) {
                RL[q] = min(RL[q], RL[q + 1]);
                if (LR) RR[q] = max(RR[q], RR[q + 1]);
            }
            // keep flags of valid windows only: bit t <-> window-end element eb + t
            {
                const uint32_t lo = e_lo > eb ? min(e_lo - eb, (uint32_t)W) : 0u;
                const uint32_t hi = e_hi > eb ? min(e_hi - eb, (uint32_t)W) : 0u;
                const uint32_t mhi = hi >= 32u ? 0xffffffffu : ((1u << hi) - 1u);
                const uint32_t mlo = lo >= 32u ? 0xffffffffu : ((1u << lo) - 1u);
                if (!SYNC && sg.first_always && e_lo >= eb && e_lo < eb + W) bf |= 1u << (e_lo - eb);
                bf &= mhi & ~mlo;
            }
            flagw[(size_t)b * NT + tid] = bf;
            cnt += __popc(bf);
        }
    }
    emit_phase(a, sg, tile, cnt, flagw, NB, es,
               [&](uint32_t q, uint32_t bit, uint32_t& jv, uint32_t& d) {
                   const uint32_t e = q * W + bit;  // window-end element
                   const uint32_t jl = e - (W - 1);
                   jv = jl - sg.has_prev;
                   const uint32_t wv = recw[((size_t)q * WQ + (bit >> 2)) * NT + tid];
                   const uint32_t lowb = (wv >> (8u * (bit & 3u))) & 0xffu;
                   d = (lowb - jl) & 0xffu;  // selected k-mer index is in [jl, jl + W)
               });
}

// ---- host side ---------------------------------------------------------------------------
// Geometry for the fast kernel; returns false when (k, w, ...) is outside its domain.
inline bool plan_fast(int sm_count, size_t smem_optin, const mz_params& p, uint64_t nwin,
                      uint32_t* NT, uint32_t* S, size_t* smem, uint32_t* num_tiles) {
    if (p.w > FAST_MAX_W) return false;
    if (p.hash_canonical && !p.strand_tiebreak) return false;  // rare combo -> generic kernel
    const char* env_nt = getenv("MZ_FAST_NT");
    const char* env_s = getenv("MZ_FAST_S");
    uint32_t nt = env_nt ? (uint32_t)atoi(env_nt) : 128u;
    if (nt != 64 && nt != 128 && nt != 256) nt = 128;
    const size_t budget = std::min<size_t>(smem_optin, 220 * 1024);
    uint32_t s;
    if (env_s) {
        s = (uint32_t)atoi(env_s);
    } else {
        // enough tiles to fill the machine a few times over, S in [64, 416]
        uint64_t target_tiles = (uint64_t)sm_count * 8;
        uint64_t want = (nwin + target_tiles * nt - 1) / (target_tiles * nt);
        s = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(want, 64), 416);
    }
    s = std::max<uint32_t>(16, (s + 15) / 16 * 16);
    while (s > 16 && fast_smem(nt, s, p.w) > budget / 3) s -= 16;
    if (fast_smem(nt, s, p.w) > budget) return false;
    if ((uint64_t)s + p.w + 2 >= 65535) return false;
    const uint64_t Tt = (uint64_t)nt * s;
    const uint64_t tiles = (nwin + Tt - 1) / Tt;
    if (tiles == 0 || tiles > 0x7fffffffull) return false;
    *NT = nt, *S = s, *smem = fast_smem(nt, s, p.w), *num_tiles = (uint32_t)tiles;
    return true;
}

template <int W, bool HC, bool LR, bool SYNC>
inline int launch_fast_inst(uint32_t NT, size_t smem, uint32_t tiles, const KArgs& a, cudaStream_t st) {
    auto kern = mz_fast_kernel<W, HC, LR, SYNC>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return MZ_ERR_CUDA;
    }
    kern<<<tiles, NT, smem, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? MZ_OK : MZ_ERR_CUDA;
}

template <int W>
inline int launch_fast_w(const mz_params& p, uint32_t NT, size_t smem, uint32_t tiles, const KArgs& a,
                         cudaStream_t st) {
    const bool sync = p.mode != MZ_MODE_MINIMIZER;
    if (p.strand_tiebreak) {
        return sync ? launch_fast_inst<W, true, true, true>(NT, smem, tiles, a, st)
                    : launch_fast_inst<W, true, true, false>(NT, smem, tiles, a, st);
    }
    return sync ? launch_fast_inst<W, false, false, true>(NT, smem, tiles, a, st)
                : launch_fast_inst<W, false, false, false>(NT, smem, tiles, a, st);
}

// defined in mz_fast_g{0..3}.cu (W = 1..8, 9..16, 17..24, 25..32)
int launch_fast_g0(const mz_params&, uint32_t, size_t, uint32_t, const KArgs&, cudaStream_t);
int launch_fast_g1(const mz_params&, uint32_t, size_t, uint32_t, const KArgs&, cudaStream_t);
int launch_fast_g2(const mz_params&, uint32_t, size_t, uint32_t, const KArgs&, cudaStream_t);
int launch_fast_g3(const mz_params&, uint32_t, size_t, uint32_t, const KArgs&, cudaStream_t);

inline int launch_fast(const mz_params& p, uint32_t NT, size_t smem, uint32_t tiles, const KArgs& a,
                       cudaStream_t st) {
    switch ((p.w - 1) / 8) {
        case 0: return launch_fast_g0(p, NT, smem, tiles, a, st);
        case 1: return launch_fast_g1(p, NT, smem, tiles, a, st);
        case 2: return launch_fast_g2(p, NT, smem, tiles, a, st);
        case 3: return launch_fast_g3(p, NT, smem, tiles, a, st);
        default: return MZ_ERR_UNSUPPORTED;
    }
}

}  // namespace mz
