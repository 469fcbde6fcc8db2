// mz_emit.cuh -- what a thread works on (Segment) and the ordered emission phase shared by the
// generic and the W-specialised kernels.
//
// Phase 1 (per kernel) leaves, per thread, a bitmask of windows that emit and the selected
// k-mer of each window.  Phase 2 (here) turns that into the reference's output order:
//   block scan of per-thread counts -> decoupled look-back for the tile's global offset ->
//   entries staged through shared memory in chunks -> coalesced stores of pos / sk / values.
// Output order = window order, which is what collect_and_dedup produces (src/collect.rs:15-76).
#pragma once
#include "mz_common.cuh"

namespace mz {

struct Segment {
    uint64_t bit0;      // bit position (in a.seq) of local base 0
    uint64_t pos_base;  // added to a local k-mer index to give the output position
    uint64_t win_base;  // window index reported for local valid window 0
    uint32_t nvalid;    // windows this thread may emit
    uint32_t has_prev;  // first computed window only seeds the dedup comparison
    uint32_t first_always;
};

// Segment of thread `tid` of tile `tile`.  Single sequence: S consecutive windows, plus one
// window on the left whose result only seeds the dedup comparison (the reference's lane-seam
// rule, src/collect.rs:252-272).  Batch mode: one read per thread.
// NT = threads per tile.
__device__ __forceinline__ Segment make_segment_nt(const KArgs& a, uint32_t tile, uint32_t tid,
                                                   uint32_t NT) {
    Segment s;
    if (a.n_reads == 0) {
        uint64_t j0 = a.wbeg + ((uint64_t)tile * NT + tid) * a.S;
        uint64_t left = j0 < a.wend ? a.wend - j0 : 0;
        s.nvalid = (uint32_t)(left < a.S ? left : a.S);
        s.has_prev = (j0 > 0 && a.mode == MODE_MINIMIZER) ? 1u : 0u;
        uint64_t s0 = j0 - s.has_prev;
        s.bit0 = (uint64_t)((int64_t)(2 * s0) + a.bitbias);
        s.pos_base = s0;
        s.win_base = j0;
        s.first_always = (j0 == 0);
    } else {
        // batch mode: thread = read, or = piece of S windows of a long read
        const uint64_t pi = (uint64_t)tile * NT + tid;
        s.has_prev = 0;
        s.first_always = 1;
        s.pos_base = 0;
        s.win_base = 0;
        s.nvalid = 0;
        s.bit0 = 0;
        if (pi < a.n_reads) {
            const uint64_t r = a.piece_read ? a.piece_read[pi] : pi;
            const uint32_t win0 = a.piece_read ? a.piece_win0[pi] : 0u;
            const uint64_t startbits = a.read_start_bp ? 2 * a.read_start_bp[r] : r * a.stride_bits;
            const uint32_t len = a.read_len_bp ? a.read_len_bp[r] : a.fixed_len_bp;
            const uint32_t nw = len >= a.l ? len - a.l + 1 : 0;
            const uint32_t left = nw > win0 ? nw - win0 : 0;
            s.nvalid = left < a.S ? left : a.S;
            s.has_prev = (win0 > 0 && a.mode == MODE_MINIMIZER) ? 1u : 0u;
            s.first_always = (win0 == 0);
            s.pos_base = win0 - s.has_prev;
            s.win_base = win0;
            s.bit0 = (uint64_t)((int64_t)(startbits + 2ull * s.pos_base) + a.bitbias);
        }
    }
    return s;
}

__device__ __forceinline__ Segment make_segment(const KArgs& a, uint32_t tile, uint32_t tid) {
    return make_segment_nt(a, tile, tid, blockDim.x);
}

// CSR end offset of a read, written by the thread that owns the read's last piece.
__device__ __forceinline__ void write_csr_offset(const KArgs& a, uint64_t pi, unsigned long long end) {
    if (pi >= a.n_reads) return;
    uint64_t r = pi;
    if (a.piece_read) {
        r = a.piece_read[pi];
        if (pi + 1 < a.n_reads && a.piece_read[pi + 1] == r) return;  // not the last piece
    }
    a.out_offsets[r + 1] = end;
    if (r == 0) a.out_offsets[0] = 0;
}

constexpr uint32_t EMIT_CHUNK = 2048;  // entries staged per round
// shared memory used by emit_phase: scratch (40 words) + list_a (u32) + list_b (u16)
constexpr uint32_t EMIT_SMEM_BYTES = 40 * 4 + EMIT_CHUNK * 4 + EMIT_CHUNK * 2;

struct EmitSmem {
    uint32_t* scratch;  // 40 words
    uint32_t* list_a;   // (owner tid << 16) | jv
    uint16_t* list_b;   // selected k-mer index - local window index
    __device__ __forceinline__ static EmitSmem carve(unsigned char* p) {
        EmitSmem e;
        e.scratch = reinterpret_cast<uint32_t*>(p);
        e.list_a = e.scratch + 40;
        e.list_b = reinterpret_cast<uint16_t*>(e.list_a + EMIT_CHUNK);
        return e;
    }
};

// Flag word q of a thread lives at flagw[q * fstride * NT + tid].
// dec(q, bit, jv, d): decode bit `bit` of this thread's flag word q into the valid-window index
// jv and d = (selected local k-mer index) - (local window index).
template <typename DecFn>
__device__ __forceinline__ void emit_phase(const KArgs& a, const Segment& sg, uint32_t tile,
                                           uint32_t cnt, const uint32_t* flagw, uint32_t fstride,
                                           uint32_t nq, const EmitSmem& es, DecFn dec) {
    const uint32_t NT = blockDim.x, tid = threadIdx.x;
    __shared__ unsigned long long s_gbase;
    uint32_t total;
    const uint32_t toff = block_exclusive_scan(cnt, es.scratch, total);
    if (tid < 32) {
        unsigned long long g = lookback_warp0(a.tile_state, tile, total);
        if (tid == 0) {
            s_gbase = g;
            if (tile == a.num_tiles - 1) {
                *a.count_out = g + total;
                if (a.h_count) *a.h_count = g + total;
            }
        }
    }
    __syncthreads();
    const unsigned long long gbase = s_gbase;
    const bool ovf = gbase + total > a.cap;
    if (ovf && tid == 0) {
        *a.overflow = 1u;
        if (a.h_overflow) *a.h_overflow = 1u;
    }
    if (a.n_reads != 0) write_csr_offset(a, (uint64_t)tile * NT + tid, gbase + toff + cnt);
    if (ovf || total == 0) return;

    const bool minim = a.mode == MODE_MINIMIZER;
    const bool canon_val = a.val_canonical != 0;
    uint32_t q = 0, produced = 0;
    uint32_t m = (cnt && nq) ? flagw[tid] : 0u;
    for (uint32_t cbase = 0; cbase < total; cbase += EMIT_CHUNK) {
        // stage this thread's entries that fall into [cbase, cbase + EMIT_CHUNK)
        while (produced < cnt && toff + produced < cbase + EMIT_CHUNK) {
            while (m == 0) {
                q++;
                m = flagw[(size_t)q * fstride * NT + tid];
            }
            uint32_t bit = (uint32_t)__ffs(m) - 1u;
            m &= m - 1u;
            uint32_t jv, d;
            dec(q, bit, jv, d);
            uint32_t slot = toff + produced - cbase;
            es.list_a[slot] = (tid << 16) | jv;
            es.list_b[slot] = (uint16_t)d;
            produced++;
        }
        __syncthreads();
        const uint32_t nent = min(EMIT_CHUNK, total - cbase);
        for (uint32_t i = tid; i < nent; i += NT) {
            const uint32_t ea = es.list_a[i];
            const uint32_t owner = ea >> 16, jv = ea & 0xffffu, d = es.list_b[i];
            const Segment og = make_segment(a, tile, owner);
            const uint32_t jl = jv + og.has_prev;
            const uint32_t local = minim ? jl + d : jl;  // local base index of the reported k/l-mer
            const unsigned long long o = gbase + cbase + i;
            a.pos[o] = (uint32_t)(og.pos_base + local);
            if (a.want_sk) a.sk[o] = (uint32_t)(og.win_base + jv);
            if (a.value_bits == 64) {
                a.val[o] = kmer_value_u64(a, og.bit0 + 2ull * local, a.val_len, canon_val);
            } else if (a.value_bits == 128) {
                uint64_t lo, hi;
                kmer_value_u128(a, og.bit0 + 2ull * local, a.val_len, canon_val, lo, hi);
                reinterpret_cast<ulonglong2*>(a.val)[o] = make_ulonglong2(lo, hi);
            }
        }
        __syncthreads();
    }
}

}  // namespace mz
