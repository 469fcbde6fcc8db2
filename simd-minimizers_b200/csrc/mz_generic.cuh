// mz_generic.cuh -- runtime-(k,w) minimizer / syncmer kernel: any k, any w that fits the
// shared-memory ring.  One thread walks S consecutive windows of the sequence (or one read in
// batch mode) with a register-resident rolling hash and a two-stacks sliding minimum whose
// ring lives in shared memory; results are recorded per window and emitted in a second,
// ordered phase (block scan + decoupled look-back).  The W-specialised fast kernel
// (mz_fast.cuh) shares phase 2 with this one.
#pragma once
#include "mz_emit.cuh"

namespace mz {

// Dynamic shared memory layout of the generic kernel (NT = blockDim.x):
//   uint2 tab[16]; uint32 misc[8]; emit staging (EMIT_SMEM_BYTES);
//   uint32 flagw[ceil(S/32)][NT]; uint16 rec[S][NT]; ring[W][NT] (uint2 if LR else uint32)
// GRING: the two-stacks ring lives in global memory (very large w) instead of shared memory;
// a template flag so that the common case compiles to LDS/STS rather than generic LD/ST.
template <bool HC, bool LR, bool GRING>
__global__ void __launch_bounds__(256) mz_generic_kernel(const KArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t NT = blockDim.x, tid = threadIdx.x;
    uint2* tab = reinterpret_cast<uint2*>(smem_raw);
    uint32_t* misc = reinterpret_cast<uint32_t*>(tab + 16);
    const EmitSmem es = EmitSmem::carve(reinterpret_cast<unsigned char*>(misc + 8));
    uint32_t* flagw = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(misc + 8) + EMIT_SMEM_BYTES);
    const uint32_t nfw = (a.S + 31) / 32;
    uint16_t* rec = reinterpret_cast<uint16_t*>(flagw + (size_t)nfw * NT);
    // ring starts 8-byte aligned: S*NT*2 bytes with NT multiple of 32 is a multiple of 8
    // very large w: the ring does not fit shared memory and lives in a per-block global region
    unsigned char* ring_raw = GRING
        ? reinterpret_cast<unsigned char*>(a.scratch + (size_t)blockIdx.x * a.scratch_words_per_block)
        : reinterpret_cast<unsigned char*>(rec + (size_t)a.S * NT);
    uint2* ring2 = reinterpret_cast<uint2*>(ring_raw);
    uint32_t* ring1 = reinterpret_cast<uint32_t*>(ring_raw);

    const uint32_t k = a.k, w = a.w, l = a.l, R = a.rot;
    if (tid < 16) {
        uint32_t in = tid & 3u, out = tid >> 2;
        // delta of one rolling step that adds base `in` and drops base `out` (k bases earlier)
        uint32_t dfw = a.f[in] ^ rotl32(a.f[out], (R * k) & 31u);
        uint32_t drc = rotl32(a.c[in], (R * (k - 1)) & 31u) ^ rotr32(a.c[out], R & 31u);
        tab[tid] = make_uint2(dfw, drc);
    }
    if (tid == 0) {
        // hash state of the virtual all-'A' k-mer that precedes every segment
        uint32_t fa = 0, ca = 0;
        for (uint32_t j = 0; j < k; j++) {
            fa ^= rotl32(a.f[0], (R * j) & 31u);
            ca ^= rotl32(a.c[0], (R * j) & 31u);
        }
        misc[1] = fa;
        misc[2] = ca;
    }
    for (;;) {  // persistent blocks: one tile per iteration
    __syncthreads();
    if (tid == 0) misc[0] = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const uint32_t tile = misc[0];
    if (tile >= a.num_tiles) break;
    const Segment sg = make_segment(a, tile, tid);

    uint32_t cnt = 0;
    if (sg.nvalid) {
        uint32_t fw = misc[1], rc = misc[2];
        BaseReader in, out;
        in.init(a, sg.bit0);
        out.init(a, sg.bit0);
        for (uint32_t s = 0; s < w; s++) {
            if (LR) ring2[s * NT + tid] = make_uint2(0xffffffffu, 0u);
            else ring1[s * NT + tid] = 0xffffffffu;
        }
        uint32_t slot = 0, preL = 0xffffffffu, preR = 0u;
        uint32_t prev = 0xffffffffu, flags = 0;
        const uint32_t nsteps = sg.nvalid + sg.has_prev + l - 1;
        const uint32_t mode = a.mode;
        // ambiguity mask (run_skip_ambiguous_windows): zrun = unambiguous bases ending at base t
        const bool amb = a.amb != nullptr;
        const uint64_t ab0 = (uint64_t)((int64_t)sg.pos_base + a.amb_bitbias);
        uint32_t zrun = 0, amw = 0;
        bool pclean = false;
        for (uint32_t t = 0; t < nsteps; t++) {
            if (amb) {
                if ((t & 31u) == 0) amw = amb_bits32(a, ab0 + t);
                zrun = (amw & 1u) ? 0u : zrun + 1u;
                amw >>= 1;
            }
            uint32_t b_in = in.next(a);
            uint32_t b_out = t >= k ? out.next(a) : 0u;
            uint2 d = tab[b_in | (b_out << 2)];
            fw = rotl32(fw, R) ^ d.x;
            uint32_t h = fw;
            if (HC) {
                rc = rotr32(rc, R) ^ d.y;
                h = fw + rc;
            }
            if (t + 1 < k) continue;
            const uint32_t e = t - (k - 1);  // local k-mer index
            const uint32_t le = (h & 0xffff0000u) | e;
            const uint32_t re = (~h & 0xffff0000u) | e;
            uint32_t sufL, sufR = 0;
            if (LR) {
                ring2[slot * NT + tid] = make_uint2(le, re);
                preL = min(preL, le);
                preR = max(preR, re);
                if (++slot == w) {
                    slot = 0;
                    uint2 sfx = ring2[(w - 1) * NT + tid];
                    for (uint32_t q = w - 1; q-- > 0;) {
                        uint2 x = ring2[q * NT + tid];
                        sfx.x = min(sfx.x, x.x);
                        sfx.y = max(sfx.y, x.y);
                        ring2[q * NT + tid] = sfx;
                    }
                    preL = le;
                    preR = re;
                }
                uint2 sf = ring2[slot * NT + tid];
                sufL = sf.x;
                sufR = sf.y;
            } else {
                ring1[slot * NT + tid] = le;
                preL = min(preL, le);
                if (++slot == w) {
                    slot = 0;
                    uint32_t sfx = ring1[(w - 1) * NT + tid];
                    for (uint32_t q = w - 1; q-- > 0;) {
                        sfx = min(sfx, ring1[q * NT + tid]);
                        ring1[q * NT + tid] = sfx;
                    }
                    preL = le;
                }
                sufL = ring1[slot * NT + tid];
            }
            if (e + 1 < w) continue;
            const uint32_t jl = e - (w - 1);  // local window index
            uint32_t sel = min(preL, sufL) & 0xffffu;
            if (LR) {
                uint32_t selR = max(preR, sufR) & 0xffffu;
                if (sel != selR) {
                    // strand rule, evaluated only when leftmost != rightmost
                    uint32_t tg = tg_count(a, sg.bit0 + 2ull * jl, l);
                    if (!(2 * tg > l)) sel = selR;
                }
            }
            const int jv = (int)jl - (int)sg.has_prev;
            bool flag;
            if (mode == MODE_MINIMIZER) flag = (jv == 0 && sg.first_always) || sel != prev;
            else if (mode == MODE_CLOSED) flag = sel == jl || sel == jl + w - 1;
            else flag = sel == jl + w / 2;
            if (amb) {
                // a window with an ambiguous base emits nothing; the first clean window after
                // one is compared against SKIPPED in the reference, i.e. always emitted
                const bool clean = zrun >= l;
                if (mode == MODE_MINIMIZER) flag = flag || !pclean;
                flag = flag && clean;
                pclean = clean;
            }
            prev = sel;
            if (jv >= 0) {
                rec[(uint32_t)jv * NT + tid] = (uint16_t)sel;
                if (flag) flags |= 1u << (jv & 31);
                if ((jv & 31) == 31) {
                    flagw[((uint32_t)jv >> 5) * NT + tid] = flags;
                    cnt += __popc(flags);
                    flags = 0;
                }
            }
        }
        if (sg.nvalid & 31u) {
            flagw[(sg.nvalid >> 5) * NT + tid] = flags;
            cnt += __popc(flags);
        }
    }
    emit_phase(a, sg, tile, cnt, flagw, 1u, (sg.nvalid + 31u) / 32u, es,
               [&](uint32_t q, uint32_t bit, uint32_t& jv, uint32_t& d) {
                   jv = q * 32u + bit;
                   d = (uint32_t)rec[jv * NT + tid] - (jv + sg.has_prev);
               });
    }  // persistent loop
}

}  // namespace mz
