// mz_api.cu -- C ABI (include/mz_b200.h) over the sm_100a minimizer kernels.
// Host side only orchestrates: shard windows over devices, copy, launch, gather.
// There is deliberately no CPU implementation here.
#include "../../include/mz_b200.h"

#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "mz_fast.cuh"
#include "mz_generic.cuh"

namespace {

thread_local std::string g_last_error;

int cuda_fail(cudaError_t e, const char* what, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s failed at mz_api.cu:%d: %s (%s)", what, line,
             cudaGetErrorString(e), cudaGetErrorName(e));
    g_last_error = buf;
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? MZ_ERR_NO_DEVICE
                                                                       : MZ_ERR_CUDA;
}
#define CK(call)                                                   \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call, __LINE__); \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    int reserve(size_t n) {
        if (n <= cap) return MZ_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 1024;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&p, n * sizeof(T));
            want = n;
        }
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc", __LINE__);
        cap = want;
        return MZ_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// grow-only pinned host buffer (bounce buffer for callers whose memory is pageable)
struct PinBuf {
    unsigned char* p = nullptr;
    size_t cap = 0;  // bytes
    int reserve(size_t n) {
        if (n <= cap) return MZ_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 4096;
        cudaError_t e = cudaHostAlloc((void**)&p, want, cudaHostAllocPortable);
        if (e != cudaSuccess) return cuda_fail(e, "cudaHostAlloc", __LINE__);
        cap = want;
        return MZ_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

// true when `ptr` is ordinary pageable host memory (cudaMemcpyAsync from/to it is staged by the
// driver at a fraction of the PCIe rate)
bool is_pageable(const void* ptr) {
    if (!ptr) return false;
    cudaPointerAttributes at{};
    cudaError_t e = cudaPointerGetAttributes(&at, ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
}

// host threads used for bounce-buffer copies and delta decoding: 16 by default, MZ_HOST_THREADS
// overrides (several processes sharing one host, e.g. one rank per GPU)
unsigned host_threads() {
    static const unsigned v = [] {
        unsigned n = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
        if (const char* e = getenv("MZ_HOST_THREADS")) n = (unsigned)std::max(1, atoi(e));
        return std::min(n, 64u);
    }();
    return v;
}

// memcpy split over a few host threads (a single thread cannot keep up with a Gen5 x16 link)
void parallel_memcpy(void* dst, const void* src, size_t bytes, unsigned max_threads = 0) {
    unsigned nt = max_threads ? std::min(max_threads, host_threads()) : host_threads();
    if (bytes < (8u << 20)) nt = 1;
    if (nt == 1) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = ((bytes + nt - 1) / nt + 4095) & ~size_t(4095);  // nt * per >= bytes
    for (unsigned i = 0; i < nt; i++) {
        const size_t b = std::min(bytes, (size_t)i * per), e = std::min(bytes, b + per);
        if (e > b) th.emplace_back([=] { memcpy((char*)dst + b, (const char*)src + b, e - b); });
    }
    for (auto& t : th) t.join();
}

struct HostScalars {  // pinned
    unsigned long long count;
    uint32_t ticket;
    uint32_t overflow;
};

struct DevState {
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    // H2D begin/end, kernel end, D2H end, D2H begin, kernel end kept until the slot's next chunk
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    DevBuf<unsigned long long> scratch;  // [0]=count [1]=ticket|overflow [2..]=tile_state
    DevBuf<uint32_t> rows;               // fast kernel per-block record rows
    DevBuf<uint8_t> in;
    DevBuf<uint8_t> ascii;               // mz_pack_ascii / mz_run_ascii staging
    DevBuf<uint8_t> amb;                 // ambiguity mask, one bit per base (skip-ambiguous runs)
    DevBuf<uint32_t> pos, sk;
    DevBuf<uint64_t> val;
    DevBuf<uint64_t> offs;  // batch CSR offsets
    DevBuf<uint64_t> rstart;
    DevBuf<uint32_t> rlen;
    DevBuf<uint32_t> pread, pwin;  // batch pieces of long reads
    HostScalars* hs = nullptr;      // mapped pinned memory: the kernels write count / overflow here
    HostScalars* hs_dev = nullptr;  // device address of hs
    PinBuf st_in, st_pos, st_sk, st_val, st_amb;  // pinned bounce buffers (pageable callers)
    PinBuf st_offs;                               // batch: chunk-local CSR offsets on their way out
    DevBuf<uint8_t> delta;                        // delta-coded pos / sk of a chunk (transfer codec)
    PinBuf st_delta;
    // on-device consumer (mz_run_bucket_stats): per-device histograms and per-chunk seam records
    DevBuf<unsigned long long> hist;
    PinBuf st_tail;
    // front-loaded uploads of the chunk pipelines (slot 0 of a device only): the device's whole
    // share of the input, copied on its own stream as early as the link allows
    DevBuf<uint8_t> in_all;
    cudaStream_t up = nullptr;
    cudaEvent_t up_t0 = nullptr, up_t1 = nullptr;
    std::vector<cudaEvent_t> up_ev;  // one per chunk, grown on demand
};

}  // namespace

constexpr int kSlots = 4;  // chunks in flight per device in the pipelined host path

struct mz_ctx {
    std::vector<DevState> devs;                 // slot 0 of every device
    std::vector<std::vector<DevState>> extra;   // slots 1..kSlots-1 of every device
    mz_timing timing{};
    // The sequence the last single-launch mz_run left in devs[0].in (mz_values reuses it instead of
    // uploading it again: Output::values_*() right after run()); cleared by every other use of the buffer.
    struct Resident {
        const uint8_t* packed = nullptr;
        uint64_t bp_offset = 0, n_bp = 0, byte_lo = 0;
        size_t nbytes = 0;
    } resident;
    DevState& slot(size_t dev, int s) { return s == 0 ? devs[dev] : extra[dev][s - 1]; }
};

namespace {

struct Plan {
    bool fast = false;
    uint32_t NT = 128, S = 32;
    size_t smem = 0;
    uint32_t num_tiles = 0;
    uint32_t grid = 0;            // generic kernel: persistent blocks
    size_t ring_words = 0;        // generic kernel, huge w: per-block ring in global memory
};

uint32_t round_up(uint32_t x, uint32_t m) { return (x + m - 1) / m * m; }

size_t generic_smem(uint32_t NT, uint32_t S, uint32_t w, bool lr) {
    size_t fixed = 16 * 8 + 8 * 4 + mz::EMIT_SMEM_BYTES;
    size_t per_thread = (size_t)((S + 31) / 32) * 4 + (size_t)S * 2 + (size_t)w * (lr ? 8 : 4);
    return fixed + per_thread * NT;
}

// Choose launch geometry for a range of `nwin` windows (single sequence) on `d`.
int plan_generic(const DevState& d, const mz_params& p, uint64_t nwin, Plan* pl) {
    const bool lr = p.strand_tiebreak != 0;
    const size_t budget = std::min<size_t>(d.smem_optin, 200 * 1024);
    uint32_t NT = 128;
    // shrink the block until a ring of w entries + 32 windows per thread fits
    while (NT >= 32 && generic_smem(NT, 32, p.w, lr) > budget) NT /= 2;
    bool global_ring = false;
    if (NT < 32) {  // w too large for a shared-memory ring: keep it in global memory (slow path)
        NT = 64;
        global_ring = true;
    }
    // Long windows: a ring of w entries per thread in shared memory leaves room for very few
    // windows per thread (w = 301: S = 32, an 11x halo, 2.8 Gbp/s).  A ring in the L2-resident
    // global scratch with 128 threads and S = 352 measured 17 Gbp/s (w = 301), 6.3 (w = 1001).
    if (p.w >= 256 && !getenv("MZ_GENERIC_SMEM_RING")) {
        NT = 128;
        global_ring = true;
    }
    if (const char* e = getenv("MZ_GENERIC_GRING")) {  // experiments: force the global ring
        NT = (uint32_t)std::max(32, atoi(e));
        global_ring = true;
    }
    // target ~4 tiles per SM, S in [32, 512], multiple of 32
    uint64_t target_tiles = (uint64_t)d.sm_count * 4;
    uint64_t s = (nwin + target_tiles * NT - 1) / (target_tiles * NT);
    uint32_t S = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(s, 32), global_ring ? 352 : 512);
    S = round_up(S, 32);
    const uint32_t wring = global_ring ? 0u : p.w;  // ring entries held in shared memory
    while (S > 32 && generic_smem(NT, S, wring, lr) > budget / 2) S -= 32;
    if (const char* e = getenv("MZ_GENERIC_S")) {
        S = round_up(std::max(32, atoi(e)), 32);
        while (S > 32 && generic_smem(NT, S, wring, lr) > budget) S -= 32;
    }
    if (S + p.w + 2 >= 65535) return MZ_ERR_UNSUPPORTED;
    pl->fast = false;
    pl->NT = NT;
    pl->S = S;
    pl->smem = generic_smem(NT, S, wring, lr);
    pl->ring_words = global_ring ? (size_t)p.w * NT * (lr ? 2 : 1) : 0;
    uint64_t T = (uint64_t)NT * S;
    uint64_t tiles = (nwin + T - 1) / T;
    if (tiles == 0 || tiles > 0x7fffffffull) return MZ_ERR_UNSUPPORTED;
    pl->num_tiles = (uint32_t)tiles;
    const uint32_t bps = global_ring ? 2u : (uint32_t)std::max<size_t>(1, std::min<size_t>(8, (220 * 1024) / (pl->smem + 1024)));
    pl->grid = (uint32_t)std::min<uint64_t>(tiles, (uint64_t)d.sm_count * bps);
    return MZ_OK;
}

void fill_hash_args(mz::KArgs& a, const mz_params& p) {
    a.k = p.k;
    a.w = p.w;
    a.l = p.k + p.w - 1;
    a.mode = p.mode;
    a.want_sk = p.want_sk ? 1u : 0u;
    a.value_bits = p.value_bits;
    a.val_len = p.mode == MZ_MODE_MINIMIZER ? p.k : p.k + p.w - 1;
    a.val_canonical = p.strand_tiebreak ? 1u : 0u;
    a.rot = p.rot;
    for (int i = 0; i < 4; i++) a.f[i] = p.f[i], a.c[i] = p.c[i];
}

template <typename K>
int set_smem(K kernel, size_t smem) {
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return MZ_OK;
}

int launch_generic(const mz_params& p, const Plan& pl, const mz::KArgs& a, cudaStream_t st) {
    const bool hc = p.hash_canonical != 0, lr = p.strand_tiebreak != 0, gr = a.scratch != nullptr;
    int rc;
#define MZ_LAUNCH_GENERIC(HC, LR, GR)                                                     \
    do {                                                                                  \
        if ((rc = set_smem(mz::mz_generic_kernel<HC, LR, GR>, pl.smem))) return rc;       \
        mz::mz_generic_kernel<HC, LR, GR><<<pl.grid, pl.NT, pl.smem, st>>>(a);            \
    } while (0)
    if (hc && lr) {
        if (gr) MZ_LAUNCH_GENERIC(true, true, true); else MZ_LAUNCH_GENERIC(true, true, false);
    } else if (hc) {
        if (gr) MZ_LAUNCH_GENERIC(true, false, true); else MZ_LAUNCH_GENERIC(true, false, false);
    } else {
        if (gr) MZ_LAUNCH_GENERIC(false, false, true); else MZ_LAUNCH_GENERIC(false, false, false);
    }
#undef MZ_LAUNCH_GENERIC
    CK(cudaGetLastError());
    return MZ_OK;
}

// Point the kernel at the mapped host scalars and clear them (the previous launch that used this
// DevState has been retired by the caller).
bool hs_by_copy() {  // experiments: MZ_HS_COPY=1 restores the D2H copy of the scalars
    static const bool v = getenv("MZ_HS_COPY") != nullptr;
    return v;
}
void attach_host_scalars(DevState& d, mz::KArgs& a) {
    if (hs_by_copy()) return;
    d.hs->count = 0;
    d.hs->ticket = 0;
    d.hs->overflow = 0;
    a.h_count = &d.hs_dev->count;
    a.h_overflow = &d.hs_dev->overflow;
}

// Enqueue one launch producing windows [wbeg, wend) on d.stream.  Input is already on the
// device (a.seq etc. filled by the caller).  Count / overflow land in d.hs; caller synchronises.
int enqueue_run(DevState& d, const mz_params& p, mz::KArgs a, uint64_t wbeg, uint64_t wend,
                uint32_t* launches) {
    Plan pl;
    mz::FastPlan fp;
    int rc = MZ_OK;
    if (mz::plan_fast(d.sm_count, p, wend - wbeg, &fp, /*allow_xw=*/a.n_reads == 0)) {
        pl.fast = true;
        pl.S = fp.S;
        pl.num_tiles = fp.num_tiles;
        if ((rc = d.rows.reserve((fp.scratch_words_per_block * 2 + fp.r1_words) * fp.grid * mz::FAST_WARPS))) return rc;
        a.scratch = d.rows.p;
        mz::fast_plan_args(fp, a);
    } else {
        rc = plan_generic(d, p, wend - wbeg, &pl);
        a.scratch = nullptr;
        if (!rc && pl.ring_words) {
            if ((rc = d.rows.reserve(pl.ring_words * pl.grid))) return rc;
            a.scratch = d.rows.p;
            a.scratch_words_per_block = pl.ring_words;
        }
    }
    if (rc) return rc;
    if ((rc = d.scratch.reserve(2 + (size_t)pl.num_tiles))) return rc;
    CK(cudaMemsetAsync(d.scratch.p, 0, (2 + (size_t)pl.num_tiles) * sizeof(unsigned long long), d.stream));
    a.wbeg = wbeg;
    a.wend = wend;
    a.S = pl.S;
    a.num_tiles = pl.num_tiles;
    a.count_out = d.scratch.p;
    a.ticket = reinterpret_cast<uint32_t*>(d.scratch.p + 1);
    a.overflow = a.ticket + 1;
    a.tile_state = d.scratch.p + 2;
    attach_host_scalars(d, a);
    if (pl.fast) rc = mz::launch_fast(p, fp.grid, a, d.stream);
    else rc = launch_generic(p, pl, a, d.stream);
    if (rc > 0 && rc != MZ_OK) {
        if (rc == MZ_ERR_CUDA && g_last_error.empty()) g_last_error = "kernel launch failed";
        return rc;
    }
    if (launches) (*launches)++;
    if (hs_by_copy()) CK(cudaMemcpyAsync(d.hs, d.scratch.p, sizeof(HostScalars), cudaMemcpyDeviceToHost, d.stream));
    return MZ_OK;
}

// Kernel time of a chunk in a pipeline: ev[1] .. ev[2] of its slot, minus the time its launch spent
// queued behind the previous chunk of the same device (another stream; every launch occupies all
// SMs, so the launches of one device run one after the other): the device was busy with THIS chunk
// from max(ev[1], kernel end of the previous chunk) on.  prev = slot of the device's previous chunk
// (its ev[5] has completed), or nullptr.
float chunk_kernel_ms(DevState& d, DevState* prev) {
    float ker = 0, since_prev = 0;
    if (cudaEventElapsedTime(&ker, d.ev[1], d.ev[2]) != cudaSuccess) {
        cudaGetLastError();
        return 0.f;
    }
    if (prev && cudaEventElapsedTime(&since_prev, prev->ev[5], d.ev[2]) == cudaSuccess) {
        if (since_prev >= 0.f && since_prev < ker) ker = since_prev;
    } else {
        cudaGetLastError();
    }
    return ker;
}

// the n-th upload-done event of a device (created on first use, kept for the context's lifetime)
int upload_event(DevState& d0, size_t n, cudaEvent_t* ev) {
    while (d0.up_ev.size() <= n) {
        cudaEvent_t e = nullptr;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        d0.up_ev.push_back(e);
    }
    *ev = d0.up_ev[n];
    return MZ_OK;
}

uint64_t estimate_capacity(const mz_params& p, uint64_t nwin) {
    double dens;
    if (p.mode == MZ_MODE_MINIMIZER) dens = 2.0 / (p.w + 1.0);
    else if (p.mode == MZ_MODE_CLOSED_SYNCMER) dens = p.w == 1 ? 1.0 : 2.0 / p.w;
    else dens = 1.0 / p.w;
    double est = nwin * dens * 1.2 + 65536.0;
    return (uint64_t)std::min<double>(est, (double)nwin);
}

cudaError_t init_devstate(DevState& d, int device) {
    d.device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d.up, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&d.up_t0);
    if (e == cudaSuccess) e = cudaEventCreate(&d.up_t1);
    for (int j = 0; j < 6 && e == cudaSuccess; j++) e = cudaEventCreate(&d.ev[j]);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&d.hs, sizeof(HostScalars), cudaHostAllocPortable | cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&d.hs_dev, d.hs, 0);
    int v = 0;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
    d.sm_count = v;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    d.smem_optin = (size_t)v;
    return e;
}

int check_values(const mz_params& p) {
    if (p.value_bits == 0) return MZ_OK;
    if (p.value_bits != 64 && p.value_bits != 128) return MZ_ERR_BAD_ARG;
    uint32_t len = p.mode == MZ_MODE_MINIMIZER ? p.k : p.k + p.w - 1;
    if (len > p.value_bits / 2) return MZ_ERR_VALUE_WIDTH;
    return MZ_OK;
}

// Fill the kernel arguments that describe a device copy of bases [blo, ...) of the sequence.
void fill_input_args(mz::KArgs& a, const mz_params& p, const uint8_t* d_in, uint64_t bp_offset,
                     uint64_t byte_lo, size_t nbytes, uint64_t nwin) {
    fill_hash_args(a, p);
    a.seq = reinterpret_cast<const uint32_t*>(d_in);
    a.bitbias = (int64_t)(2 * bp_offset) - (int64_t)(8 * byte_lo);
    a.seq_nwords = (nbytes + 3) / 4;
    a.nwin = nwin;
}

// Ambiguity mask of a skip-ambiguous run: one bit per base, base g -> bit (off + g) of `bits`.
struct AmbSrc {
    const uint8_t* bits = nullptr;
    uint64_t off = 0;
};

// Byte range of the mask that covers bases [blo, bhi), start aligned down to 4 bytes.
void amb_range(const AmbSrc& am, uint64_t blo, uint64_t bhi, uint64_t* byte_lo, size_t* nbytes) {
    *byte_lo = ((am.off + blo) / 8) & ~uint64_t(3);
    *nbytes = (size_t)((am.off + bhi + 7) / 8 - *byte_lo);
}

void fill_amb_args(mz::KArgs& a, const AmbSrc& am, const uint8_t* d_amb, uint64_t byte_lo, size_t nbytes) {
    a.amb = reinterpret_cast<const uint32_t*>(d_amb);
    a.amb_bitbias = (int64_t)am.off - (int64_t)(8 * byte_lo);
    a.amb_nwords = (nbytes + 3) / 4;
}

// H2D of the mask bytes covering bases [blo, bhi) into d.amb (+ kernel arguments).
int upload_amb(DevState& d, const AmbSrc& am, uint64_t blo, uint64_t bhi, uint64_t* byte_lo, size_t* nbytes) {
    amb_range(am, blo, bhi, byte_lo, nbytes);
    int rc;
    if ((rc = d.amb.reserve(*nbytes + 64))) return rc;
    const uint8_t* src = am.bits + *byte_lo;
    if (is_pageable(am.bits) && *nbytes >= (1u << 20)) {
        if ((rc = d.st_amb.reserve(*nbytes))) return rc;
        parallel_memcpy(d.st_amb.p, src, *nbytes);
        src = d.st_amb.p;
    }
    // the word holding the last mask bits may extend past the copied bytes: zero it first
    CK(cudaMemsetAsync(d.amb.p + (*nbytes & ~size_t(3)), 0, 8, d.stream));
    CK(cudaMemcpyAsync(d.amb.p, src, *nbytes, cudaMemcpyHostToDevice, d.stream));
    return MZ_OK;
}

// ---- transfer codec for minimizer positions / super-k-mer starts --------------------------------
// Entry i+1 is emitted at the first window j whose selection differs from window j-1's, which was
// entry i's position p_i in [j-1, j+w-2]; the new position lies in [j, j+w-1], so
// p_{i+1} - p_i is in [-(w-2), w] whatever the tie rule does, and a run of windows that all select
// the same k-mer is at most w long, so consecutive super-k-mer starts differ by 1..w.  For w <= 127 a chunk's u32 array
// crosses PCIe as one signed byte per entry plus an absolute u32 every 256 entries: 1.02 instead of
// 4 bytes per entry (C2: 3.72 -> 2.8 GB of D2H per run).  The host adds the deltas up again while it
// writes the caller's array (16 threads, a few ms per 3.1 Gbp, hidden behind the next chunk).
constexpr uint32_t kDeltaBlock = 256;
__global__ void mz_delta_encode_kernel(const uint32_t* __restrict__ v, uint64_t n,
                                       int8_t* __restrict__ delta, uint32_t* __restrict__ base) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if ((i & (kDeltaBlock - 1)) == 0) {
        base[i / kDeltaBlock] = v[i];
        delta[i] = 0;
    } else {
        delta[i] = (int8_t)(int32_t)(v[i] - v[i - 1]);
    }
}
}  // namespace
namespace mz {
void delta_decode_blocks(const int8_t* delta, const uint32_t* base, uint64_t n, uint64_t b0, uint64_t b1, uint32_t* out);  // mz_host_codec.cpp
}
namespace {
size_t delta_bytes(uint64_t n) {  // [deltas, padded to 4][bases]
    return (size_t)((n + 3) & ~uint64_t(3)) + (size_t)((n + kDeltaBlock - 1) / kDeltaBlock) * 4;
}
void delta_decode(const unsigned char* enc, uint64_t n, uint32_t* out, unsigned max_threads = 0) {
    const int8_t* delta = reinterpret_cast<const int8_t*>(enc);
    const uint32_t* base = reinterpret_cast<const uint32_t*>(enc + ((n + 3) & ~uint64_t(3)));
    const uint64_t nblk = (n + kDeltaBlock - 1) / kDeltaBlock;
    unsigned nt = max_threads ? std::min(max_threads, host_threads()) : host_threads();
    if (nblk < 64) nt = 1;
    auto work = [=](uint64_t b0, uint64_t b1) { mz::delta_decode_blocks(delta, base, n, b0, b1, out); };
    if (nt == 1) {
        work(0, nblk);
        return;
    }
    std::vector<std::thread> th;
    const uint64_t per = (nblk + nt - 1) / nt;
    for (unsigned t = 0; t < nt; t++) {
        const uint64_t b0 = std::min(nblk, (uint64_t)t * per), b1 = std::min(nblk, b0 + per);
        if (b1 > b0) th.emplace_back(work, b0, b1);
    }
    for (auto& t : th) t.join();
}

// Large inputs, any number of devices: windows are cut into chunks; chunk c runs on device
// c % ndev in stream slot (c / ndev) % kSlots of that device, so that on every device the H2D copy
// of one chunk, the kernel of the next and the D2H copy of an earlier one overlap, and all devices
// (one PCIe link each) move data at the same time.  Chunks are seams like any other shard (one
// extra window on the left); because they are dealt out in order, the output offset of chunk c
// (the sum of the counts of all chunks before it) is known as soon as the kernels of chunks
// 0..c have finished, whichever devices they ran on, and every D2H copy lands directly at its
// final place in the caller's arrays: one ordered, globally indexed output, no second pass.
int run_pipelined(mz_ctx* ctx, size_t ndev, const mz_params& p, const uint8_t* packed, uint64_t bp_offset,
                  uint64_t n_bp, const AmbSrc& am, mz_out* out) {
    const uint32_t l = p.k + p.w - 1;
    const uint64_t nwin = n_bp - l + 1;
    const uint32_t vw = p.value_bits / 64;
    // >= 24 chunks, >= 6 per device, but never tiny ones (launch + copy latency)
    uint64_t chunk = std::max<uint64_t>(ndev == 1 ? (1ull << 25) : (1ull << 22),
                                        (nwin + std::max<uint64_t>(24, 6 * ndev) - 1) / std::max<uint64_t>(24, 6 * ndev));
    const char* env = getenv("MZ_CHUNK_WINDOWS");
    if (env) chunk = std::max<uint64_t>(1, strtoull(env, nullptr, 10));
    const uint64_t nchunks = (nwin + chunk - 1) / chunk;
    const uint64_t ND = ndev;
    struct Job {
        uint64_t wb, we, cap, byte_lo, count = 0, out_off = 0, abyte_lo = 0;
        size_t nbytes, anbytes = 0;
        bool staged = false;
        float decode_ms = 0, t_out_ms = 0;  // MZ_DEBUG_PIPE: decode time, D2H completion since the call began
        const uint8_t* d_in = nullptr;  // this chunk's bytes on its device
        cudaEvent_t up_done = nullptr;  // front-loaded upload: the copy of this chunk has landed
    };
    // Pageable caller memory goes through pinned bounce buffers + multi-threaded memcpy; pinned
    // (mz_host_alloc / cudaHostRegister'ed) memory is used directly.
    const bool page_in = is_pageable(packed);
    const bool page_out = is_pageable(out->pos) || (p.want_sk && is_pageable(out->sk)) || (vw && is_pageable(out->val));
    // positions (and super-k-mer starts) of plain minimizer runs cross PCIe delta-coded.  With
    // several devices the links add up but the host cores that decode do not: the codec is used
    // while the decode rate (MZ_DELTA_MAX_DEVICES, default 2) keeps up with the links.
    static const uint64_t delta_max_dev = getenv("MZ_DELTA_MAX_DEVICES") ? strtoull(getenv("MZ_DELTA_MAX_DEVICES"), nullptr, 10) : 2;
    const bool delta = p.mode == MZ_MODE_MINIMIZER && p.w <= 127 && !am.bits && !getenv("MZ_NO_POS_DELTA") && ND <= delta_max_dev;
    // host threads per decode / copy-out job: the jobs of all devices run side by side
    const unsigned job_threads = std::min(8u, std::max(2u, host_threads() / (unsigned)std::min<uint64_t>(ND, 4)));
    std::vector<Job> jobs(nchunks);
    uint64_t total = 0;
    bool too_small = false;
    int rc;
    double dbg_sync_ms = 0;  // MZ_DEBUG_PIPE: time the calling thread waits for kernels
    const auto t_start = std::chrono::steady_clock::now();
    std::vector<float> dev_h2d(ND, 0.f), dev_ker(ND, 0.f), dev_d2h(ND, 0.f);

    // D2H time of the chunk that used a slot last (its copies have completed when this is called)
    std::vector<char> d2h_pending(ndev * kSlots, 0);
    auto collect_d2h = [&](size_t di, int sl) {
        if (!d2h_pending[di * kSlots + sl]) return;
        d2h_pending[di * kSlots + sl] = 0;
        DevState& d = ctx->slot(di, sl);
        float ms = 0;
        if (cudaEventElapsedTime(&ms, d.ev[4], d.ev[3]) == cudaSuccess) dev_d2h[di] += ms;
        else cudaGetLastError();
    };
    auto dev_of = [&](uint64_t c) { return (size_t)(c % ND); };
    auto slot_of = [&](uint64_t c) { return (int)((c / ND) % kSlots); };
    // error paths return early: no stream may still be writing the caller's arrays, and no
    // joinable thread may be destroyed (declared first, so it runs after the helpers are joined)
    struct SyncAll {
        mz_ctx* ctx;
        size_t ndev;
        ~SyncAll() {
            for (size_t i = 0; i < ndev; i++)
                for (int sl = 0; sl < kSlots; sl++) {
                    DevState& d = ctx->slot(i, sl);
                    if (cudaSetDevice(d.device) == cudaSuccess) {
                        cudaStreamSynchronize(d.stream);
                        if (sl == 0) cudaStreamSynchronize(d.up);
                    }
                }
            cudaGetLastError();
            cudaSetDevice(ctx->devs[0].device);
        }
    } sync_all{ctx, ndev};

    // geometry of every chunk
    for (uint64_t c = 0; c < nchunks; c++) {
        Job& j = jobs[c];
        j.wb = c * chunk;
        j.we = std::min<uint64_t>(j.wb + chunk, nwin);
        j.cap = estimate_capacity(p, j.we - j.wb);
        const uint64_t blo = j.wb > 0 ? j.wb - 1 : 0;
        j.byte_lo = (bp_offset + blo) / 4 & ~uint64_t(3);
        j.nbytes = (bp_offset + j.we + l - 1 + 3) / 4 - j.byte_lo;
    }
    // Front-loaded uploads (pinned callers): a copy that runs while the other direction is busy
    // slows both down (measured: 55 GB/s one way, 44 + 44 GB/s both ways), and the outputs are
    // 3-5x the input.  So the whole input goes up first, on a stream of its own, as fast as the
    // link takes it; the kernels wait for their chunk's event, and most of the D2H traffic then has
    // the link to itself.  (Pageable input is staged chunk by chunk by the calling thread instead.)
    const bool front = !page_in && !getenv("MZ_NO_FRONT_UPLOAD");
    if (front) {
        std::vector<size_t> share(ND, 0);
        for (uint64_t c = 0; c < nchunks; c++) share[dev_of(c)] += (jobs[c].nbytes + 64 + 255) & ~size_t(255);
        for (size_t i = 0; i < ndev; i++) {
            CK(cudaSetDevice(ctx->devs[i].device));
            if ((rc = ctx->devs[i].in_all.reserve(share[i] + 256))) return rc;
            CK(cudaEventRecord(ctx->devs[i].up_t0, ctx->devs[i].up));
        }
        std::vector<size_t> off(ND, 0), nth(ND, 0);
        for (uint64_t c = 0; c < nchunks; c++) {
            const size_t di = dev_of(c);
            DevState& d0 = ctx->devs[di];
            Job& j = jobs[c];
            CK(cudaSetDevice(d0.device));
            uint8_t* dst = d0.in_all.p + off[di];
            off[di] += (j.nbytes + 64 + 255) & ~size_t(255);
            CK(cudaMemcpyAsync(dst, packed + j.byte_lo, j.nbytes, cudaMemcpyHostToDevice, d0.up));
            if ((rc = upload_event(d0, nth[di]++, &j.up_done))) return rc;
            CK(cudaEventRecord(j.up_done, d0.up));
            j.d_in = dst;
        }
        for (size_t i = 0; i < ndev; i++) {
            CK(cudaSetDevice(ctx->devs[i].device));
            CK(cudaEventRecord(ctx->devs[i].up_t1, ctx->devs[i].up));
        }
    }

    auto issue = [&](uint64_t c) -> int {
        DevState& d = ctx->slot(dev_of(c), slot_of(c));
        Job& j = jobs[c];
        CK(cudaSetDevice(d.device));
        const uint64_t blo = j.wb > 0 ? j.wb - 1 : 0;
        int r;
        if (!front && (r = d.in.reserve(j.nbytes + 64))) return r;
        if ((r = d.pos.reserve(j.cap))) return r;
        if (p.want_sk && (r = d.sk.reserve(j.cap))) return r;
        if (vw && (r = d.val.reserve(j.cap * vw))) return r;
        const uint8_t* src = packed + j.byte_lo;
        if (page_in) {
            if ((r = d.st_in.reserve(j.nbytes))) return r;
            parallel_memcpy(d.st_in.p, src, j.nbytes);
            src = d.st_in.p;
        }
        CK(cudaEventRecord(d.ev[0], d.stream));
        if (front) {
            CK(cudaStreamWaitEvent(d.stream, j.up_done, 0));
        } else {
            CK(cudaMemcpyAsync(d.in.p, src, j.nbytes, cudaMemcpyHostToDevice, d.stream));
            j.d_in = d.in.p;
        }
        if (am.bits && (r = upload_amb(d, am, blo, j.we + l - 1, &j.abyte_lo, &j.anbytes))) return r;
        CK(cudaEventRecord(d.ev[1], d.stream));
        mz::KArgs a{};
        fill_input_args(a, p, j.d_in, bp_offset, j.byte_lo, j.nbytes, nwin);
        if (am.bits) fill_amb_args(a, am, d.amb.p, j.abyte_lo, j.anbytes);
        a.pos = d.pos.p, a.sk = d.sk.p, a.val = d.val.p, a.cap = j.cap;
        if ((r = enqueue_run(d, p, a, j.wb, j.we, &ctx->timing.kernel_launches))) return r;
        CK(cudaEventRecord(d.ev[2], d.stream));
        CK(cudaEventRecord(d.ev[5], d.stream));
        return MZ_OK;
    };
    // D2H of chunk c enqueued -> a helper thread waits for it and writes the caller's arrays (delta
    // decode / copy out of the bounce buffers), so the calling thread keeps the pipelines fed
    const size_t nhelp = ndev * kSlots;
    std::vector<std::thread> helper(nhelp);
    std::vector<int> helper_rc(nhelp, 0);
    struct JoinAll {
        std::vector<std::thread>& h;
        ~JoinAll() {
            for (auto& t : h)
                if (t.joinable()) t.join();
        }
    } join_all{helper};
    auto join_helper = [&](size_t hi) -> int {
        if (helper[hi].joinable()) helper[hi].join();
        const int r = helper_rc[hi];
        helper_rc[hi] = 0;
        return r;
    };
    auto finish = [&](uint64_t c) -> int {
        const size_t hi = dev_of(c) * kSlots + slot_of(c);
        DevState& d = ctx->slot(dev_of(c), slot_of(c));
        Job& j = jobs[c];
        if (!(j.staged || delta) || !j.count || too_small) return MZ_OK;
        int r = join_helper(hi);
        if (r) return r;
        const bool want_sk = p.want_sk != 0;
        int* const hrc = &helper_rc[hi];
        auto work = [&d, &j, hrc, out, vw, delta, want_sk, job_threads, t_start]() {
            if (cudaSetDevice(d.device) != cudaSuccess || cudaEventSynchronize(d.ev[3]) != cudaSuccess) {
                *hrc = MZ_ERR_CUDA;
                return;
            }
            const auto t0 = std::chrono::steady_clock::now();
            j.t_out_ms = std::chrono::duration<float, std::milli>(t0 - t_start).count();
            if (delta) {
                delta_decode(d.st_delta.p, j.count, out->pos + j.out_off, job_threads);
                if (want_sk) delta_decode(d.st_delta.p + delta_bytes(j.count), j.count, out->sk + j.out_off, job_threads);
            } else {
                parallel_memcpy(out->pos + j.out_off, d.st_pos.p, j.count * 4, job_threads);
                if (want_sk) parallel_memcpy(out->sk + j.out_off, d.st_sk.p, j.count * 4, job_threads);
            }
            if (vw && j.staged) parallel_memcpy(out->val + j.out_off * vw, d.st_val.p, j.count * 8 * vw, job_threads);
            j.decode_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        };
        try {
            helper[hi] = std::thread(work);
        } catch (...) {  // no thread to be had: do it here (nothing may unwind across the C ABI)
            work();
            r = helper_rc[hi];
            helper_rc[hi] = 0;
            if (r) return r;
        }
        return MZ_OK;
    };
    auto retire = [&](uint64_t c) -> int {
        const size_t di = dev_of(c);
        DevState& d = ctx->slot(di, slot_of(c));
        Job& j = jobs[c];
        CK(cudaSetDevice(d.device));
        {
            const auto ts0 = std::chrono::steady_clock::now();
            CK(cudaEventSynchronize(d.ev[2]));
            dbg_sync_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ts0).count();
        }
        float h2d = 0;
        cudaEventElapsedTime(&h2d, d.ev[0], d.ev[1]);
        dev_h2d[di] += h2d;
        dev_ker[di] += chunk_kernel_ms(d, c >= ND ? &ctx->slot(di, slot_of(c - ND)) : nullptr);
        collect_d2h(di, slot_of(c));  // the slot's previous chunk (stream order: it is out)
        uint64_t count = d.hs->count;
        if (d.hs->overflow) {  // capacity estimate too small: redo this chunk with the exact size
            int r;
            j.cap = count;
            if ((r = d.pos.reserve(j.cap))) return r;
            if (p.want_sk && (r = d.sk.reserve(j.cap))) return r;
            if (vw && (r = d.val.reserve(j.cap * vw))) return r;
            mz::KArgs a{};
            fill_input_args(a, p, j.d_in, bp_offset, j.byte_lo, j.nbytes, nwin);
            if (am.bits) fill_amb_args(a, am, d.amb.p, j.abyte_lo, j.anbytes);
            a.pos = d.pos.p, a.sk = d.sk.p, a.val = d.val.p, a.cap = j.cap;
            if ((r = enqueue_run(d, p, a, j.wb, j.we, &ctx->timing.kernel_launches))) return r;
            CK(cudaStreamSynchronize(d.stream));
            if (d.hs->overflow) {
                g_last_error = "internal: exact-capacity re-run overflowed";
                return MZ_ERR_CUDA;
            }
            count = d.hs->count;
        }
        if (total + count > out->capacity) too_small = true;
        j.count = count;
        j.out_off = total;
        if (!too_small && count) {
            // the slot's bounce buffers are about to be overwritten: its previous chunk must be out
            if (int r = join_helper(di * kSlots + slot_of(c))) return r;
            uint32_t *hpos = out->pos + total, *hsk = p.want_sk ? out->sk + total : nullptr;
            uint64_t* hval = vw ? out->val + total * vw : nullptr;
            if (page_out) {
                int r;
                if (!delta && (r = d.st_pos.reserve(count * 4))) return r;
                if (!delta && p.want_sk && (r = d.st_sk.reserve(count * 4))) return r;
                if (vw && (r = d.st_val.reserve(count * 8 * vw))) return r;
                hpos = reinterpret_cast<uint32_t*>(d.st_pos.p);
                hsk = reinterpret_cast<uint32_t*>(d.st_sk.p);
                hval = reinterpret_cast<uint64_t*>(d.st_val.p);
                j.staged = true;
            }
            CK(cudaEventRecord(d.ev[4], d.stream));
            d2h_pending[di * kSlots + slot_of(c)] = 1;
            if (delta) {
                int r;
                const size_t nb = delta_bytes(count), narr = p.want_sk ? 2 : 1;
                if ((r = d.delta.reserve(nb * narr))) return r;
                if ((r = d.st_delta.reserve(nb * narr))) return r;
                const unsigned nt = 256, nblk = (unsigned)((count + nt - 1) / nt);
                const size_t boff = (size_t)((count + 3) & ~uint64_t(3));
                mz_delta_encode_kernel<<<nblk, nt, 0, d.stream>>>(d.pos.p, count, reinterpret_cast<int8_t*>(d.delta.p),
                                                                 reinterpret_cast<uint32_t*>(d.delta.p + boff));
                if (p.want_sk)
                    mz_delta_encode_kernel<<<nblk, nt, 0, d.stream>>>(d.sk.p, count, reinterpret_cast<int8_t*>(d.delta.p + nb),
                                                                     reinterpret_cast<uint32_t*>(d.delta.p + nb + boff));
                CK(cudaGetLastError());
                ctx->timing.kernel_launches += (uint32_t)narr;
                CK(cudaMemcpyAsync(d.st_delta.p, d.delta.p, nb * narr, cudaMemcpyDeviceToHost, d.stream));
            } else {
                CK(cudaMemcpyAsync(hpos, d.pos.p, count * 4, cudaMemcpyDeviceToHost, d.stream));
                if (p.want_sk) CK(cudaMemcpyAsync(hsk, d.sk.p, count * 4, cudaMemcpyDeviceToHost, d.stream));
            }
            if (vw) CK(cudaMemcpyAsync(hval, d.val.p, count * 8 * vw, cudaMemcpyDeviceToHost, d.stream));
            CK(cudaEventRecord(d.ev[3], d.stream));
        }
        total += count;
        return MZ_OK;
    };
    // chunks issued ahead of the one being retired: two per device (H2D + kernel of c, c - ndev
    // in flight while the D2H of c - 2 ndev runs), three without the helper stage
    const bool helpers = page_out || delta;
    const uint64_t R = (helpers ? kSlots - 2 : kSlots - 1) * ND, F = R + ND;
    for (uint64_t c = 0; c < nchunks + F; c++) {
        if (c < nchunks && (rc = issue(c))) return rc;
        if (c >= R && c - R < nchunks && (rc = retire(c - R))) return rc;
        if (helpers && c >= F && c - F < nchunks && (rc = finish(c - F))) return rc;
    }
    for (size_t i = 0; i < ndev; i++)
        for (int sl = 0; sl < kSlots; sl++) {
            DevState& d = ctx->slot(i, sl);
            CK(cudaSetDevice(d.device));
            CK(cudaStreamSynchronize(d.stream));
        }
    for (size_t hi = 0; hi < nhelp; hi++)
        if (int r = join_helper(hi)) return r;
    for (size_t hi = 0; hi < nhelp; hi++) collect_d2h(hi / kSlots, (int)(hi % kSlots));
    if (front)
        for (size_t i = 0; i < ndev; i++) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, ctx->devs[i].up_t0, ctx->devs[i].up_t1) == cudaSuccess) dev_h2d[i] = ms;
            else cudaGetLastError();
        }
    for (size_t i = 0; i < ndev; i++) {  // per-phase times: the busiest device
        ctx->timing.h2d_ms = std::max(ctx->timing.h2d_ms, dev_h2d[i]);
        ctx->timing.kernel_ms = std::max(ctx->timing.kernel_ms, dev_ker[i]);
        ctx->timing.d2h_ms = std::max(ctx->timing.d2h_ms, dev_d2h[i]);
    }
    CK(cudaSetDevice(ctx->devs[0].device));
    if (getenv("MZ_DEBUG_PIPE")) {
        const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
        double dec = 0;
        for (const Job& j : jobs) dec += j.decode_ms;
        fprintf(stderr, "[mz pipeline] devices %zu chunks %llu (%llu windows each) delta %d: wall %.1f ms, calling thread "
                        "waited %.1f ms for kernels, helper threads %.1f ms (%u threads per job)\n",
                ndev, (unsigned long long)nchunks, (unsigned long long)chunk, (int)delta, wall, dbg_sync_ms, dec, job_threads);
        for (size_t i = 0; i < ndev; i++)
            fprintf(stderr, "[mz pipeline]   device %d: h2d %.2f ms, kernels %.2f ms, d2h %.2f ms (sums over its chunks)\n",
                    ctx->devs[i].device, dev_h2d[i], dev_ker[i], dev_d2h[i]);
        if (atoi(getenv("MZ_DEBUG_PIPE")) > 1) {
            fprintf(stderr, "[mz pipeline]   D2H of chunk c complete at (ms):");
            for (const Job& j : jobs) fprintf(stderr, " %.1f", j.t_out_ms);
            fprintf(stderr, "\n[mz pipeline]   decode of chunk c took (ms):");
            for (const Job& j : jobs) fprintf(stderr, " %.1f", j.decode_ms);
            fprintf(stderr, "\n");
        }
    }
    out->count = total;
    return too_small ? MZ_ERR_CAPACITY : MZ_OK;
}

// Chunk-local CSR offsets -> global ones (mz_run_batch): the base of a chunk is the number of entries
// of all chunks before it, known once their kernels have finished.  Done on the device so that the
// offsets can be copied straight into the caller's array (a host loop over 200 M offsets was a
// third of config 5's end-to-end time).
__global__ void mz_rebase_offsets_kernel(unsigned long long* __restrict__ offs, uint64_t n, unsigned long long base) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) offs[i] += base;
}

// Output::values_u64 / values_u128 (src/lib.rs:598-629): one thread per position.
__global__ void mz_values_kernel(mz::KArgs a, const uint32_t* __restrict__ pos, uint64_t n, uint64_t n_bp,
                                 uint64_t* __restrict__ val, uint32_t* __restrict__ bad) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t p = pos[i];
    if (p + a.val_len > n_bp) {  // the reference's read_kmer asserts on this
        *bad = 1u;
        return;
    }
    const uint64_t bit = (uint64_t)((int64_t)(2 * p) + a.bitbias);
    if (a.value_bits == 64) {
        __stcs(reinterpret_cast<unsigned long long*>(val) + i,
               (unsigned long long)mz::kmer_value_u64(a, bit, a.val_len, a.val_canonical != 0));
    } else {
        uint64_t lo, hi;
        mz::kmer_value_u128(a, bit, a.val_len, a.val_canonical != 0, lo, hi);
        reinterpret_cast<ulonglong2*>(val)[i] = make_ulonglong2(lo, hi);
    }
}

// ---- first on-device consumer: super-k-mers bucketed by their minimizer -------------
// A super-k-mer is a maximal run of windows with the same minimizer (bench/src/minimizer.rs:3-36,
// Problem C): entry i of a .super_kmers() run starts at window sk[i] and ends where entry i+1
// starts.  Downstream tools (k-mer counters, partitioned assemblers, sparse dictionaries) shard
// super-k-mers by minimizer: bucket = floor(mix64(value) * n_buckets / 2^64), mix64 = the
// splitmix64 finaliser, value = the canonical k-mer of the minimizer (values_u64).  The kernel
// below consumes a chunk's (sk, val) arrays where the minimizer kernel left them, in HBM, and
// adds to per-device histograms: super-k-mers per bucket and windows per bucket.
__host__ __device__ inline uint64_t mz_mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
struct ChunkTail {  // what the host needs to stitch the chunk seams
    unsigned long long count;  // entries of the chunk
    uint32_t first_sk;         // first window that starts a super-k-mer inside the chunk
    uint32_t last_bucket;      // bucket of the chunk's last super-k-mer (it runs on into the next chunk)
};
constexpr uint32_t kBucketMax = 16384;  // two u32 counters per bucket in shared memory
__global__ void __launch_bounds__(1024, 1)
mz_bucket_kernel(const uint32_t* __restrict__ sk, const uint64_t* __restrict__ val, const unsigned long long* count_ptr,
                 const uint32_t* overflow, uint32_t wend, uint32_t nb, unsigned long long* g_cnt,
                 unsigned long long* g_win, ChunkTail* tail) {
    extern __shared__ uint32_t hs[];  // [nb] super-k-mers, [nb] windows
    if (*overflow) return;            // the chunk is re-run with the exact capacity first
    const unsigned long long n = *count_ptr;
    for (uint32_t i = threadIdx.x; i < 2 * nb; i += blockDim.x) hs[i] = 0;
    __syncthreads();
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t b = (uint32_t)__umul64hi(mz_mix64(val[i]), (uint64_t)nb);
        const uint32_t next = i + 1 < n ? sk[i + 1] : wend;
        atomicAdd(&hs[b], 1u);
        atomicAdd(&hs[nb + b], next - sk[i]);
        if (i + 1 == n) tail->last_bucket = b;
        if (i == 0) tail->first_sk = sk[0];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) tail->count = n;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) {
        if (hs[i]) atomicAdd(&g_cnt[i], (unsigned long long)hs[i]);
        if (hs[nb + i]) atomicAdd(&g_win[i], (unsigned long long)hs[nb + i]);
    }
}

// Chunk pipeline of the consumer: front-loaded uploads, minimizer kernel, bucket kernel; only the
// seam records (16 bytes per chunk) and, at the end, the histograms cross the bus.
int run_bucket_stats(mz_ctx* ctx, const mz_params& p, const uint8_t* packed, uint64_t bp_offset, uint64_t n_bp,
                     uint32_t nb, uint64_t* cnt_out, uint64_t* win_out, uint64_t* total_out) {
    const uint32_t l = p.k + p.w - 1;
    const uint64_t nwin = n_bp - l + 1;
    const size_t ndev = ctx->devs.size();
    const uint64_t ND = ndev;
    uint64_t chunk = std::max<uint64_t>(1ull << 22, (nwin + std::max<uint64_t>(24, 6 * ndev) - 1) / std::max<uint64_t>(24, 6 * ndev));
    if (const char* e = getenv("MZ_CHUNK_WINDOWS")) chunk = std::max<uint64_t>(1, strtoull(e, nullptr, 10));
    const uint64_t nchunks = (nwin + chunk - 1) / chunk;
    struct Job {
        uint64_t wb, we, cap, byte_lo;
        size_t nbytes;
        const uint8_t* d_in = nullptr;
        cudaEvent_t up_done = nullptr;
    };
    std::vector<Job> jobs(nchunks);
    auto dev_of = [&](uint64_t c) { return (size_t)(c % ND); };
    auto slot_of = [&](uint64_t c) { return (int)((c / ND) % kSlots); };
    struct SyncAll {
        mz_ctx* ctx;
        ~SyncAll() {
            for (size_t i = 0; i < ctx->devs.size(); i++)
                for (int sl = 0; sl < kSlots; sl++) {
                    DevState& d = ctx->slot(i, sl);
                    if (cudaSetDevice(d.device) == cudaSuccess) {
                        cudaStreamSynchronize(d.stream);
                        if (sl == 0) cudaStreamSynchronize(d.up);
                    }
                }
            cudaGetLastError();
            cudaSetDevice(ctx->devs[0].device);
        }
    } sync_all{ctx};
    int rc;
    const bool page_in = is_pageable(packed);
    std::vector<size_t> share(ND, 0), off(ND, 0), nth(ND, 0);
    for (uint64_t c = 0; c < nchunks; c++) {
        Job& j = jobs[c];
        j.wb = c * chunk;
        j.we = std::min<uint64_t>(j.wb + chunk, nwin);
        j.cap = estimate_capacity(p, j.we - j.wb);
        const uint64_t blo = j.wb > 0 ? j.wb - 1 : 0;
        j.byte_lo = (bp_offset + blo) / 4 & ~uint64_t(3);
        j.nbytes = (bp_offset + j.we + l - 1 + 3) / 4 - j.byte_lo;
        share[dev_of(c)] += (j.nbytes + 64 + 255) & ~size_t(255);
    }
    for (size_t i = 0; i < ndev; i++) {
        DevState& d0 = ctx->devs[i];
        CK(cudaSetDevice(d0.device));
        if ((rc = d0.in_all.reserve(share[i] + 256))) return rc;
        if ((rc = d0.hist.reserve(2 * (size_t)nb))) return rc;
        if ((rc = d0.st_tail.reserve(sizeof(ChunkTail) * ((nchunks + ND - 1) / ND + 1)))) return rc;
        memset(d0.st_tail.p, 0, sizeof(ChunkTail) * ((nchunks + ND - 1) / ND + 1));
        CK(cudaMemsetAsync(d0.hist.p, 0, 2 * (size_t)nb * 8, d0.up));
        CK(cudaEventRecord(d0.up_t0, d0.up));
    }
    // uploads: pinned input goes up in one sweep; pageable input chunk by chunk through the bounce buffer
    for (uint64_t c = 0; c < nchunks; c++) {
        const size_t di = dev_of(c);
        DevState& d0 = ctx->devs[di];
        Job& j = jobs[c];
        CK(cudaSetDevice(d0.device));
        uint8_t* dst = d0.in_all.p + off[di];
        off[di] += (j.nbytes + 64 + 255) & ~size_t(255);
        const uint8_t* src = packed + j.byte_lo;
        if (page_in) {
            CK(cudaStreamSynchronize(d0.up));  // the bounce buffer is free again
            if ((rc = d0.st_in.reserve(j.nbytes))) return rc;
            parallel_memcpy(d0.st_in.p, src, j.nbytes);
            src = d0.st_in.p;
        }
        CK(cudaMemcpyAsync(dst, src, j.nbytes, cudaMemcpyHostToDevice, d0.up));
        if ((rc = upload_event(d0, nth[di]++, &j.up_done))) return rc;
        CK(cudaEventRecord(j.up_done, d0.up));
        j.d_in = dst;
    }
    for (size_t i = 0; i < ndev; i++) {
        CK(cudaSetDevice(ctx->devs[i].device));
        CK(cudaEventRecord(ctx->devs[i].up_t1, ctx->devs[i].up));
    }
    auto tail_of = [&](uint64_t c) { return reinterpret_cast<ChunkTail*>(ctx->devs[dev_of(c)].st_tail.p) + c / ND; };
    auto launch = [&](uint64_t c) -> int {
        DevState& d = ctx->slot(dev_of(c), slot_of(c));
        DevState& d0 = ctx->devs[dev_of(c)];
        Job& j = jobs[c];
        int r;
        if ((r = d.pos.reserve(j.cap))) return r;
        if ((r = d.sk.reserve(j.cap))) return r;
        if ((r = d.val.reserve(j.cap))) return r;
        mz::KArgs a{};
        fill_input_args(a, p, j.d_in, bp_offset, j.byte_lo, j.nbytes, nwin);
        a.pos = d.pos.p, a.sk = d.sk.p, a.val = d.val.p, a.cap = j.cap;
        if ((r = enqueue_run(d, p, a, j.wb, j.we, &ctx->timing.kernel_launches))) return r;
        const size_t smem = 2 * (size_t)nb * 4;
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(mz_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ChunkTail* dtail = nullptr;
        CK(cudaHostGetDevicePointer((void**)&dtail, tail_of(c), 0));
        mz_bucket_kernel<<<d.sm_count, 1024, smem, d.stream>>>(d.sk.p, d.val.p, d.scratch.p, reinterpret_cast<uint32_t*>(d.scratch.p + 1) + 1,
                                                             (uint32_t)j.we, nb, d0.hist.p, d0.hist.p + nb, dtail);
        CK(cudaGetLastError());
        ctx->timing.kernel_launches++;
        return MZ_OK;
    };
    auto issue = [&](uint64_t c) -> int {
        DevState& d = ctx->slot(dev_of(c), slot_of(c));
        CK(cudaSetDevice(d.device));
        CK(cudaStreamWaitEvent(d.stream, jobs[c].up_done, 0));
        CK(cudaStreamWaitEvent(d.stream, ctx->devs[dev_of(c)].up_t0, 0));  // histogram zeroed
        CK(cudaEventRecord(d.ev[1], d.stream));
        int r = launch(c);
        if (r) return r;
        CK(cudaEventRecord(d.ev[2], d.stream));
        return MZ_OK;
    };
    std::vector<float> dev_ker(ND, 0.f);
    auto retire = [&](uint64_t c) -> int {
        DevState& d = ctx->slot(dev_of(c), slot_of(c));
        CK(cudaSetDevice(d.device));
        CK(cudaStreamSynchronize(d.stream));
        float ker = 0;
        cudaEventElapsedTime(&ker, d.ev[1], d.ev[2]);
        dev_ker[dev_of(c)] += ker;
        if (d.hs->overflow) {  // capacity estimate too small: redo the chunk with the exact size
            jobs[c].cap = d.hs->count;
            int r = launch(c);
            if (r) return r;
            CK(cudaStreamSynchronize(d.stream));
            if (d.hs->overflow) {
                g_last_error = "internal: exact-capacity re-run overflowed";
                return MZ_ERR_CUDA;
            }
        }
        return MZ_OK;
    };
    const uint64_t R = (kSlots - 1) * ND;
    for (uint64_t c = 0; c < nchunks + R; c++) {
        if (c < nchunks && (rc = issue(c))) return rc;
        if (c >= R && (rc = retire(c - R))) return rc;
    }
    // histograms of all devices + seam stitching on the host
    std::vector<unsigned long long> h(2 * (size_t)nb);
    for (uint32_t b = 0; b < nb; b++) cnt_out[b] = win_out[b] = 0;
    for (size_t i = 0; i < ndev; i++) {
        DevState& d0 = ctx->devs[i];
        CK(cudaSetDevice(d0.device));
        CK(cudaMemcpy(h.data(), d0.hist.p, h.size() * 8, cudaMemcpyDeviceToHost));
        for (uint32_t b = 0; b < nb; b++) cnt_out[b] += h[b], win_out[b] += h[nb + b];
        float ms = 0;
        if (cudaEventElapsedTime(&ms, d0.up_t0, d0.up_t1) == cudaSuccess) ctx->timing.h2d_ms = std::max(ctx->timing.h2d_ms, ms);
        ctx->timing.kernel_ms = std::max(ctx->timing.kernel_ms, dev_ker[i]);
    }
    // A chunk's last super-k-mer was closed at the chunk's last window; it really ends where the
    // next chunk's first super-k-mer starts (chunks without any entry lie entirely inside it).
    uint64_t total = 0;
    bool have_prev = false;
    uint32_t prev_bucket = 0;
    for (uint64_t c = 0; c < nchunks; c++) {
        const ChunkTail& t = *tail_of(c);
        const uint64_t lead = (t.count ? t.first_sk : jobs[c].we) - jobs[c].wb;
        if (have_prev) win_out[prev_bucket] += lead;
        if (t.count) have_prev = true, prev_bucket = t.last_bucket;
        total += t.count;
    }
    *total_out = total;
    CK(cudaSetDevice(ctx->devs[0].device));
    return MZ_OK;
}

// ASCII -> 2-bit packing, (c >> 1) & 3 per character; one thread per 16 characters.
__global__ void mz_pack_ascii_kernel(const uint8_t* __restrict__ ascii, uint64_t n,
                                     uint32_t* __restrict__ out, uint64_t nwords) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nwords) return;
    const uint64_t base = i * 16;
    uint32_t w = 0;
    if (base + 16 <= n) {
        const uint4 v = *reinterpret_cast<const uint4*>(ascii + base);
        const uint32_t q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t codes = (q[j] >> 1) & 0x03030303u;
            // gather the four 2-bit codes into one byte (multiply-shift, no carries collide)
            w |= ((codes * 0x01041040u) >> 24) << (8 * j);
        }
    } else {
        for (uint64_t j = 0; base + j < n; j++) w |= (uint32_t)((ascii[base + j] >> 1) & 3u) << (2 * j);
    }
    out[i] = w;
}

// Ambiguity mask of ASCII text: bit = 1 for everything except ACGTacgt; one thread per 32 chars.
__global__ void mz_amb_ascii_kernel(const uint8_t* __restrict__ ascii, uint64_t n,
                                    uint32_t* __restrict__ out, uint64_t nwords) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nwords) return;
    const uint64_t base = i * 32;
    uint32_t m = 0;
    if (base + 32 <= n) {
        const uint4 v0 = *reinterpret_cast<const uint4*>(ascii + base);
        const uint4 v1 = *reinterpret_cast<const uint4*>(ascii + base + 16);
        const uint32_t q[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t u = q[j] & 0xDFDFDFDFu;  // upper case
            const uint32_t ok = __vcmpeq4(u, 0x41414141u) | __vcmpeq4(u, 0x43434343u) |
                                __vcmpeq4(u, 0x47474747u) | __vcmpeq4(u, 0x54545454u);
            // one bit per byte (multiply-shift gathers bits 0, 8, 16, 24)
            m |= (((~ok & 0x01010101u) * 0x01020408u) >> 24 & 0xfu) << (4 * j);
        }
    } else {
        for (uint64_t j = 0; base + j < n; j++) {
            const uint8_t u = ascii[base + j] & 0xDFu;
            if (!(u == 'A' || u == 'C' || u == 'G' || u == 'T')) m |= 1u << j;
        }
    }
    out[i] = m;
}

// d.ascii (already on the device) -> ambiguity mask words in d.amb
int amb_on_device(DevState& d, uint64_t n, size_t* nbytes_out, uint32_t* launches) {
    const uint64_t nwords = (n + 31) / 32;
    int rc;
    if ((rc = d.amb.reserve(nwords * 4 + 64))) return rc;
    if (nwords) {
        const uint32_t nt = 256;
        mz_amb_ascii_kernel<<<(unsigned)((nwords + nt - 1) / nt), nt, 0, d.stream>>>(
            d.ascii.p, n, reinterpret_cast<uint32_t*>(d.amb.p), nwords);
        CK(cudaGetLastError());
        if (launches) (*launches)++;
    }
    *nbytes_out = (size_t)nwords * 4;
    return MZ_OK;
}

// ascii (host) -> d.ascii -> packed words in d.in; returns packed byte count
int pack_on_device(DevState& d, const char* ascii, uint64_t n, size_t* nbytes_out, uint32_t* launches) {
    const uint64_t nwords = (n + 15) / 16;
    int rc;
    if ((rc = d.ascii.reserve(n + 16))) return rc;
    if ((rc = d.in.reserve(nwords * 4 + 64))) return rc;
    CK(cudaMemcpyAsync(d.ascii.p, ascii, n, cudaMemcpyHostToDevice, d.stream));
    if (nwords) {
        const uint32_t nt = 256;
        mz_pack_ascii_kernel<<<(unsigned)((nwords + nt - 1) / nt), nt, 0, d.stream>>>(
            d.ascii.p, n, reinterpret_cast<uint32_t*>(d.in.p), nwords);
        CK(cudaGetLastError());
        if (launches) (*launches)++;
    }
    *nbytes_out = (size_t)nwords * 4;
    return MZ_OK;
}

}  // namespace

extern "C" {

uint32_t mz_abi_version(void) { return MZ_ABI_VERSION; }

const char* mz_strerror(int code) {
    switch (code) {
        case MZ_OK: return "ok";
        case MZ_ERR_BAD_ARG: return "bad argument";
        case MZ_ERR_W_RANGE: return "w must satisfy 0 < w < 2^15 (sliding_min is not tested for windows of length > 2^15)";
        case MZ_ERR_TOO_LONG: return "sliding_min returns 32bit indices. Try splitting the input into 4GB chunks first.";
        case MZ_ERR_EVEN_L: return "Window length l=k+w-1 must be odd to guarantee canonicality";
        case MZ_ERR_OPEN_EVEN_W: return "Open syncmers require odd window size, so that there is a unique middle element.";
        case MZ_ERR_NOT_CANONICAL: return "canonical minimizers need a canonical hasher (hasher.is_canonical())";
        case MZ_ERR_VALUE_WIDTH: return "k-mer (or l-mer for syncmers) does not fit the requested value width";
        case MZ_ERR_CAPACITY: return "output capacity too small; mz_out.count holds the required size";
        case MZ_ERR_UNSUPPORTED: return "parameter combination not supported by the B200 path";
        case MZ_ERR_NO_DEVICE: return "no usable CUDA device (this library has no CPU fallback)";
        case MZ_ERR_CUDA: return "CUDA error (see mz_last_error)";
        case MZ_ERR_NOMEM: return "out of memory";
        default: return "unknown error";
    }
}

const char* mz_last_error(void) { return g_last_error.c_str(); }

int mz_device_count(int* n) {
    if (!n) return MZ_ERR_BAD_ARG;
    *n = 0;
    cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *n = 0;
        return cuda_fail(e, "cudaGetDeviceCount", __LINE__) == MZ_ERR_CUDA ? MZ_ERR_NO_DEVICE
                                                                         : MZ_ERR_NO_DEVICE;
    }
    return *n > 0 ? MZ_OK : MZ_ERR_NO_DEVICE;
}

int mz_params_set_nthash(mz_params* p, uint32_t hash_canonical) {
    if (!p) return MZ_ERR_BAD_ARG;
    // seq-hash 0.2.0 NtHasher: low halves of the ntHash seeds, indexed by packed code A,C,T,G
    static const uint32_t F[4] = {0x95c60474u, 0x62a02b4cu, 0x82572324u, 0x4be24456u};
    for (int b = 0; b < 4; b++) p->f[b] = F[b], p->c[b] = F[b ^ 2];
    p->rot = 7;
    p->hash_canonical = hash_canonical ? 1u : 0u;
    return MZ_OK;
}

int mz_params_set_mulhash(mz_params* p, uint32_t hash_canonical) {
    if (!p) return MZ_ERR_BAD_ARG;
    const uint32_t C = 0x27220a95u;  // low half of the FxHash multiplier
    for (uint32_t b = 0; b < 4; b++) p->f[b] = b * C, p->c[b] = (b ^ 2u) * C;
    p->rot = 7;
    p->hash_canonical = hash_canonical ? 1u : 0u;
    return MZ_OK;
}

static int params_init(mz_params* p, uint32_t k, uint32_t w, uint32_t mode, uint32_t canonical) {
    if (!p) return MZ_ERR_BAD_ARG;
    memset(p, 0, sizeof *p);
    p->k = k;
    p->w = w;
    p->mode = mode;
    p->strand_tiebreak = canonical ? 1u : 0u;
    return MZ_OK;
}

int mz_params_nthash(mz_params* p, uint32_t k, uint32_t w, uint32_t mode, uint32_t canonical) {
    int rc = params_init(p, k, w, mode, canonical);
    return rc ? rc : mz_params_set_nthash(p, canonical);
}

int mz_params_mulhash(mz_params* p, uint32_t k, uint32_t w, uint32_t mode, uint32_t canonical) {
    int rc = params_init(p, k, w, mode, canonical);
    return rc ? rc : mz_params_set_mulhash(p, canonical);
}

int mz_params_validate(const mz_params* p, uint64_t n_bp) {
    if (!p || p->k == 0 || p->mode > 2 || p->rot > 31) return MZ_ERR_BAD_ARG;
    if (p->w == 0 || p->w >= (1u << 15)) return MZ_ERR_W_RANGE;
    if (n_bp >= (1ull << 32)) return MZ_ERR_TOO_LONG;
    if (p->strand_tiebreak && ((p->k + p->w - 1) & 1u) == 0) return MZ_ERR_EVEN_L;
    if (p->strand_tiebreak && !p->hash_canonical) return MZ_ERR_NOT_CANONICAL;
    if (p->mode == MZ_MODE_OPEN_SYNCMER && (p->w & 1u) == 0) return MZ_ERR_OPEN_EVEN_W;
    if (p->want_sk && p->mode != MZ_MODE_MINIMIZER) return MZ_ERR_BAD_ARG;
    return check_values(*p);
}

int mz_ctx_create(const int* device_ids, int n_devices, mz_ctx** out) {
    if (!out || n_devices < 0) return MZ_ERR_BAD_ARG;
    *out = nullptr;
    int ndev = 0;
    int rc = mz_device_count(&ndev);
    if (rc) return rc;
    std::vector<int> ids;
    if (device_ids == nullptr || n_devices == 0) {
        int cur = 0;
        CK(cudaGetDevice(&cur));
        ids.push_back(cur);
    } else {
        for (int i = 0; i < n_devices; i++) {
            if (device_ids[i] < 0 || device_ids[i] >= ndev) return MZ_ERR_NO_DEVICE;
            ids.push_back(device_ids[i]);
        }
    }
    mz_ctx* ctx = new mz_ctx();
    ctx->devs.resize(ids.size());
    ctx->extra.resize(ids.size());
    for (size_t i = 0; i < ids.size(); i++) {
        ctx->extra[i].resize(kSlots - 1);
        for (int sl = 0; sl < kSlots; sl++) {
            cudaError_t e = init_devstate(ctx->slot(i, sl), ids[i]);
            if (e != cudaSuccess) {
                int code = cuda_fail(e, "context setup", __LINE__);
                mz_ctx_destroy(ctx);
                return code;
            }
        }
    }
    *out = ctx;
    return MZ_OK;
}

void mz_ctx_destroy(mz_ctx* ctx) {
    if (!ctx) return;
    std::vector<DevState*> all;
    for (DevState& d : ctx->devs) all.push_back(&d);
    for (auto& v : ctx->extra)
        for (DevState& d : v) all.push_back(&d);
    for (DevState* dp : all) {
        DevState& d = *dp;
        cudaSetDevice(d.device);
        if (d.stream) cudaStreamSynchronize(d.stream);
        d.scratch.release(), d.rows.release(), d.ascii.release(), d.in.release(), d.pos.release(), d.sk.release(), d.val.release();
        d.offs.release(), d.rstart.release(), d.rlen.release(), d.pread.release(), d.pwin.release();
        d.st_in.release(), d.st_pos.release(), d.st_sk.release(), d.st_val.release(), d.st_amb.release(), d.st_offs.release(), d.st_delta.release(), d.delta.release();
        d.amb.release();
        d.in_all.release();
        d.hist.release();
        d.st_tail.release();
        for (auto& e : d.up_ev)
            if (e) cudaEventDestroy(e);
        if (d.up_t0) cudaEventDestroy(d.up_t0);
        if (d.up_t1) cudaEventDestroy(d.up_t1);
        if (d.up) cudaStreamDestroy(d.up);
        for (auto& e : d.ev)
            if (e) cudaEventDestroy(e);
        if (d.hs) cudaFreeHost(d.hs);
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    delete ctx;
}

int mz_ctx_device_count(const mz_ctx* ctx) { return ctx ? (int)ctx->devs.size() : 0; }

int mz_host_alloc(void** p, size_t bytes) {
    if (!p) return MZ_ERR_BAD_ARG;
    *p = nullptr;
    // Large buffers are spread page by page over all NUMA nodes of the host: the devices of a
    // multi-GPU context hang off different sockets, and a buffer that lives on one node makes
    // the links of the other socket's devices cross the inter-socket fabric (measured on a
    // 2-GPU box: 52.9 GB/s D2H for one device, 83 GB/s -- not 106 -- for two).  The pages are
    // placed when cudaHostAlloc pins them, so the policy only has to hold across that call.
    const bool spread = bytes >= (64u << 20) && !getenv("MZ_NO_NUMA_INTERLEAVE");
    bool policy_set = false;
#ifdef SYS_set_mempolicy
    if (spread) {
        unsigned long mask = 0;  // online nodes, e.g. "0-1" or "0,2-3"
        if (FILE* f = fopen("/sys/devices/system/node/online", "r")) {
            int a = 0, b = 0;
            char sep = 0;
            while (fscanf(f, "%d", &a) == 1) {
                b = a;
                if (fscanf(f, "%c", &sep) == 1 && sep == '-') {
                    if (fscanf(f, "%d", &b) != 1) b = a;
                    if (fscanf(f, "%c", &sep) != 1) sep = 0;
                }
                for (int i = a; i <= b && i < 64; i++) mask |= 1ul << i;
                if (sep != ',') break;
            }
            fclose(f);
        }
        if (mask & (mask - 1))  // at least two nodes
            policy_set = syscall(SYS_set_mempolicy, 3 /* MPOL_INTERLEAVE */, &mask, 64ul) == 0;
    }
#endif
    const cudaError_t e = cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocPortable);
#ifdef SYS_set_mempolicy
    if (policy_set) syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0ul);
#endif
    if (e != cudaSuccess) return cuda_fail(e, "cudaHostAlloc", __LINE__);
    return MZ_OK;
}

void mz_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int mz_alu_probe_device(int device, mz_alu_result* res);  // mz_probe.cu
int mz_alu_probe(mz_ctx* ctx, int dev_index, mz_alu_result* res) {
    if (!ctx || !res || dev_index < 0 || dev_index >= (int)ctx->devs.size()) return MZ_ERR_BAD_ARG;
    const int rc = mz_alu_probe_device(ctx->devs[dev_index].device, res);
    cudaSetDevice(ctx->devs[0].device);
    return rc;
}

int mz_params_set_tables(mz_params* p, const uint32_t f[4], const uint32_t c[4], uint32_t rot,
                         uint32_t hash_canonical) {
    if (!p || !f || !c || rot > 31) return MZ_ERR_BAD_ARG;
    for (int b = 0; b < 4; b++) p->f[b] = f[b], p->c[b] = c[b];
    p->rot = rot;
    p->hash_canonical = hash_canonical ? 1u : 0u;
    return MZ_OK;
}

int mz_run_bucket_stats(mz_ctx* ctx, const mz_params* p, const uint8_t* packed, uint64_t bp_offset, uint64_t n_bp,
                        uint32_t n_buckets, uint64_t* superkmers_out, uint64_t* windows_out, uint64_t* n_minimizers) {
    if (!ctx || !p || !superkmers_out || !windows_out || n_buckets == 0) return MZ_ERR_BAD_ARG;
    if (n_buckets > kBucketMax) return MZ_ERR_UNSUPPORTED;
    mz_params q = *p;
    q.want_sk = 1, q.value_bits = 64;
    if (q.mode != MZ_MODE_MINIMIZER) return MZ_ERR_BAD_ARG;  // super-k-mers exist for minimizers only (src/lib.rs:339)
    int rc = mz_params_validate(&q, n_bp);
    if (rc) return rc;
    for (uint32_t b = 0; b < n_buckets; b++) superkmers_out[b] = windows_out[b] = 0;
    if (n_minimizers) *n_minimizers = 0;
    const uint32_t l = q.k + q.w - 1;
    if (n_bp < l) return MZ_OK;
    if (!packed) return MZ_ERR_BAD_ARG;
    ctx->resident = mz_ctx::Resident{};
    ctx->timing = mz_timing{};
    const auto t0 = std::chrono::steady_clock::now();
    uint64_t total = 0;
    rc = run_bucket_stats(ctx, q, packed, bp_offset, n_bp, n_buckets, superkmers_out, windows_out, &total);
    ctx->timing.total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (n_minimizers) *n_minimizers = total;
    return rc;
}

int mz_values(mz_ctx* ctx, const mz_params* p, const uint8_t* packed, uint64_t bp_offset, uint64_t n_bp,
              const uint32_t* pos, uint64_t n_pos, uint32_t value_bits, uint64_t* val_out) {
    if (!ctx || !p || (value_bits != 64 && value_bits != 128)) return MZ_ERR_BAD_ARG;
    mz_params q = *p;
    q.value_bits = value_bits;
    int rc = mz_params_validate(&q, n_bp);
    if (rc) return rc;
    if (n_pos == 0) return MZ_OK;
    if (!packed || !pos || !val_out) return MZ_ERR_BAD_ARG;
    DevState& d = ctx->devs[0];
    CK(cudaSetDevice(d.device));
    ctx->timing = mz_timing{};
    const uint32_t vw = value_bits / 64;
    const uint64_t byte_lo = bp_offset / 4 & ~uint64_t(3);
    const size_t nbytes = (size_t)((bp_offset + n_bp + 3) / 4 - byte_lo);
    const mz_ctx::Resident& rs = ctx->resident;
    const bool have = rs.packed == packed && rs.bp_offset == bp_offset && rs.n_bp == n_bp && rs.byte_lo == byte_lo &&
                      rs.nbytes == nbytes && d.in.cap >= nbytes;
    CK(cudaEventRecord(d.ev[0], d.stream));
    if (!have) {
        ctx->resident = mz_ctx::Resident{};
        if ((rc = d.in.reserve(nbytes + 64))) return rc;
        CK(cudaMemcpyAsync(d.in.p, packed + byte_lo, nbytes, cudaMemcpyHostToDevice, d.stream));
    }
    if ((rc = d.pos.reserve(n_pos))) return rc;
    if ((rc = d.val.reserve(n_pos * vw))) return rc;
    if ((rc = d.scratch.reserve(4))) return rc;
    CK(cudaMemcpyAsync(d.pos.p, pos, n_pos * 4, cudaMemcpyHostToDevice, d.stream));
    CK(cudaMemsetAsync(d.scratch.p, 0, 8, d.stream));
    CK(cudaEventRecord(d.ev[1], d.stream));
    mz::KArgs a{};
    fill_input_args(a, q, d.in.p, bp_offset, byte_lo, nbytes, 0);
    const unsigned nt = 256;
    mz_values_kernel<<<(unsigned)((n_pos + nt - 1) / nt), nt, 0, d.stream>>>(a, d.pos.p, n_pos, n_bp, d.val.p,
                                                                         reinterpret_cast<uint32_t*>(d.scratch.p));
    CK(cudaGetLastError());
    ctx->timing.kernel_launches++;
    CK(cudaEventRecord(d.ev[2], d.stream));
    uint32_t bad = 0;
    CK(cudaMemcpyAsync(&bad, d.scratch.p, 4, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaMemcpyAsync(val_out, d.val.p, n_pos * 8 * vw, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaEventRecord(d.ev[3], d.stream));
    CK(cudaStreamSynchronize(d.stream));
    cudaEventElapsedTime(&ctx->timing.h2d_ms, d.ev[0], d.ev[1]);
    cudaEventElapsedTime(&ctx->timing.kernel_ms, d.ev[1], d.ev[2]);
    cudaEventElapsedTime(&ctx->timing.d2h_ms, d.ev[2], d.ev[3]);
    cudaEventElapsedTime(&ctx->timing.total_ms, d.ev[0], d.ev[3]);
    if (!have) {  // the sequence is resident now
        ctx->resident.packed = packed, ctx->resident.bp_offset = bp_offset, ctx->resident.n_bp = n_bp;
        ctx->resident.byte_lo = byte_lo, ctx->resident.nbytes = nbytes;
    }
    return bad ? MZ_ERR_BAD_ARG : MZ_OK;
}

// Host <-> device copy rate of the context's devices, all at once: the ceiling of every
// end-to-end number (one PCIe link per device, shared host memory / root complex).
int mz_pcie_probe(mz_ctx* ctx, uint64_t bytes_per_device, uint32_t reps, mz_pcie_result* res) {
    if (!ctx || !res || bytes_per_device == 0 || reps == 0) return MZ_ERR_BAD_ARG;
    memset(res, 0, sizeof *res);
    const size_t ndev = ctx->devs.size();
    struct Bufs {
        void *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
    };
    std::vector<Bufs> b(ndev);
    int rc = MZ_OK;
    auto cleanup = [&] {
        for (size_t i = 0; i < ndev; i++) {
            cudaSetDevice(ctx->devs[i].device);
            if (b[i].h_in) cudaFreeHost(b[i].h_in);
            if (b[i].h_out) cudaFreeHost(b[i].h_out);
            if (b[i].d_in) cudaFree(b[i].d_in);
            if (b[i].d_out) cudaFree(b[i].d_out);
        }
        cudaSetDevice(ctx->devs[0].device);
    };
    for (size_t i = 0; i < ndev && rc == MZ_OK; i++) {
        cudaError_t e = cudaSetDevice(ctx->devs[i].device);
        if (e == cudaSuccess && mz_host_alloc(&b[i].h_in, bytes_per_device) != MZ_OK) e = cudaErrorMemoryAllocation;
        if (e == cudaSuccess && mz_host_alloc(&b[i].h_out, bytes_per_device) != MZ_OK) e = cudaErrorMemoryAllocation;
        if (e == cudaSuccess) e = cudaMalloc(&b[i].d_in, bytes_per_device);
        if (e == cudaSuccess) e = cudaMalloc(&b[i].d_out, bytes_per_device);
        if (e == cudaSuccess) e = cudaMemset(b[i].d_out, 1, bytes_per_device);
        if (e != cudaSuccess) rc = cuda_fail(e, "mz_pcie_probe setup", __LINE__);
        else {
            memset(b[i].h_in, 1, bytes_per_device);   // touch the pages
            memset(b[i].h_out, 1, bytes_per_device);
        }
    }
    // mode 0: H2D only, 1: D2H only, 2: both directions at once (slot 0 / slot 1 streams)
    auto run = [&](int mode, double* h2d_gbs, double* d2h_gbs) -> int {
        for (int pass = 0; pass < 2; pass++) {  // pass 0 warms up
            const uint32_t n = pass ? reps : 1;
            const auto t0 = std::chrono::steady_clock::now();
            for (uint32_t r = 0; r < n; r++)
                for (size_t i = 0; i < ndev; i++) {
                    CK(cudaSetDevice(ctx->devs[i].device));
                    if (mode != 1) CK(cudaMemcpyAsync(b[i].d_in, b[i].h_in, bytes_per_device, cudaMemcpyHostToDevice, ctx->slot(i, 0).stream));
                    if (mode != 0) CK(cudaMemcpyAsync(b[i].h_out, b[i].d_out, bytes_per_device, cudaMemcpyDeviceToHost, ctx->slot(i, 1).stream));
                }
            for (size_t i = 0; i < ndev; i++) {
                CK(cudaSetDevice(ctx->devs[i].device));
                CK(cudaStreamSynchronize(ctx->slot(i, 0).stream));
                CK(cudaStreamSynchronize(ctx->slot(i, 1).stream));
            }
            const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            const double gbs = (double)bytes_per_device * n * ndev / s / 1e9;
            if (pass) {
                if (mode != 1) *h2d_gbs = gbs;
                if (mode != 0) *d2h_gbs = gbs;
            }
        }
        return MZ_OK;
    };
    if (rc == MZ_OK) rc = run(0, &res->h2d_gbs, nullptr);
    if (rc == MZ_OK) rc = run(1, nullptr, &res->d2h_gbs);
    if (rc == MZ_OK) rc = run(2, &res->bidir_h2d_gbs, &res->bidir_d2h_gbs);
    res->n_devices = (uint32_t)ndev;
    cleanup();
    return rc;
}

int mz_last_timing(const mz_ctx* ctx, mz_timing* t) {
    if (!ctx || !t) return MZ_ERR_BAD_ARG;
    *t = ctx->timing;
    return MZ_OK;
}

static int run_device_impl(mz_ctx* ctx, int dev_index, const mz_params* p, const void* d_packed,
                           uint64_t bp_offset, uint64_t n_bp, const void* d_amb, uint64_t amb_bit_offset,
                           bool skip_ambiguous, uint64_t win_begin, uint64_t win_end, mz_out* out) {
    if (!ctx || !p || !out || dev_index < 0 || dev_index >= (int)ctx->devs.size()) return MZ_ERR_BAD_ARG;
    int rc = mz_params_validate(p, n_bp);
    if (rc) return rc;
    if (skip_ambiguous) {
        if (!p->strand_tiebreak) return MZ_ERR_NOT_CANONICAL;  // Builder<'h, true, ..> only, src/lib.rs:451
        if (p->want_sk) return MZ_ERR_BAD_ARG;                 // ... with SuperKmers = ()
        if (!d_amb && n_bp) return MZ_ERR_BAD_ARG;
    }
    out->count = 0;
    const uint32_t l = p->k + p->w - 1;
    if (n_bp < l) return MZ_OK;
    const uint64_t nwin = n_bp - l + 1;
    if (win_end == 0 || win_end > nwin) win_end = nwin;
    if (win_begin >= win_end) return MZ_OK;
    if (!d_packed || !out->pos || (p->want_sk && !out->sk) || (p->value_bits && !out->val)) return MZ_ERR_BAD_ARG;
    // the kernels store whole elements: u32 positions, u64 values, (lo, hi) pairs as one 16-byte store
    auto misaligned = [](const void* q, uintptr_t al) { return (reinterpret_cast<uintptr_t>(q) & (al - 1)) != 0; };
    if (misaligned(out->pos, 4) || (p->want_sk && misaligned(out->sk, 4)) ||
        (p->value_bits && misaligned(out->val, p->value_bits == 128 ? 16 : 8)))
        return MZ_ERR_BAD_ARG;

    DevState& d = ctx->devs[dev_index];
    CK(cudaSetDevice(d.device));
    mz::KArgs a{};
    fill_hash_args(a, *p);
    // align the word pointer down to 4 bytes and fold the remainder into the bit bias
    uintptr_t addr = reinterpret_cast<uintptr_t>(d_packed);
    uintptr_t aligned = addr & ~uintptr_t(3);
    a.seq = reinterpret_cast<const uint32_t*>(aligned);
    a.bitbias = (int64_t)(8 * (addr - aligned) + 2 * bp_offset);
    a.seq_nwords = ((uint64_t)a.bitbias + 2 * n_bp + 31) / 32;
    a.nwin = nwin;
    if (skip_ambiguous) {
        const uintptr_t aaddr = reinterpret_cast<uintptr_t>(d_amb), aal = aaddr & ~uintptr_t(3);
        a.amb = reinterpret_cast<const uint32_t*>(aal);
        a.amb_bitbias = (int64_t)(8 * (aaddr - aal) + amb_bit_offset);
        a.amb_nwords = ((uint64_t)a.amb_bitbias + n_bp + 31) / 32;
    }
    a.pos = out->pos;
    a.sk = out->sk;
    a.val = out->val;
    a.cap = out->capacity;
    ctx->timing = mz_timing{};
    CK(cudaEventRecord(d.ev[0], d.stream));
    rc = enqueue_run(d, *p, a, win_begin, win_end, &ctx->timing.kernel_launches);
    if (rc) return rc;
    CK(cudaEventRecord(d.ev[1], d.stream));
    CK(cudaStreamSynchronize(d.stream));
    CK(cudaEventElapsedTime(&ctx->timing.kernel_ms, d.ev[0], d.ev[1]));
    ctx->timing.total_ms = ctx->timing.kernel_ms;
    out->count = d.hs->count;
    return d.hs->overflow ? MZ_ERR_CAPACITY : MZ_OK;
}

int mz_run_device(mz_ctx* ctx, int dev_index, const mz_params* p, const void* d_packed,
                  uint64_t bp_offset, uint64_t n_bp, uint64_t win_begin, uint64_t win_end,
                  mz_out* out) {
    return run_device_impl(ctx, dev_index, p, d_packed, bp_offset, n_bp, nullptr, 0, false, win_begin,
                           win_end, out);
}

int mz_run_device_skip_ambiguous(mz_ctx* ctx, int dev_index, const mz_params* p, const void* d_packed,
                                 uint64_t bp_offset, uint64_t n_bp, const void* d_ambiguous,
                                 uint64_t amb_bit_offset, uint64_t win_begin, uint64_t win_end,
                                 mz_out* out) {
    return run_device_impl(ctx, dev_index, p, d_packed, bp_offset, n_bp, d_ambiguous, amb_bit_offset,
                           true, win_begin, win_end, out);
}

static int run_host_impl(mz_ctx* ctx, const mz_params* p, const uint8_t* packed, uint64_t bp_offset,
                         uint64_t n_bp, const AmbSrc& am, bool skip_ambiguous, mz_out* out) {
    if (!ctx || !p || !out) return MZ_ERR_BAD_ARG;
    int rc = mz_params_validate(p, n_bp);
    if (rc) return rc;
    if (skip_ambiguous) {
        if (!p->strand_tiebreak) return MZ_ERR_NOT_CANONICAL;
        if (p->want_sk) return MZ_ERR_BAD_ARG;
        if (!am.bits && n_bp) return MZ_ERR_BAD_ARG;
    }
    out->count = 0;
    ctx->resident = mz_ctx::Resident{};
    const uint32_t l = p->k + p->w - 1;
    if (n_bp < l) return MZ_OK;
    if (!packed || !out->pos || (p->want_sk && !out->sk) || (p->value_bits && !out->val)) return MZ_ERR_BAD_ARG;
    const uint64_t nwin = n_bp - l + 1;
    const size_t ndev = ctx->devs.size();
    const uint32_t vw = p->value_bits / 64;  // u64 words per value
    ctx->timing = mz_timing{};
    // Large inputs flow through the chunk pipeline (all devices of the context); small ones are
    // one launch on the first device (sharding a few million windows costs more than it saves).
    uint64_t pipe_min = ndev == 1 ? (1ull << 26) : (1ull << 23);
    if (const char* e = getenv("MZ_PIPELINE_MIN_WINDOWS")) pipe_min = strtoull(e, nullptr, 10);
    if (nwin >= pipe_min && !getenv("MZ_NO_PIPELINE")) {
        const auto t0 = std::chrono::steady_clock::now();
        rc = run_pipelined(ctx, ndev, *p, packed, bp_offset, n_bp, am, out);
        ctx->timing.total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return rc;
    }

    DevState& d = ctx->devs[0];
    CK(cudaSetDevice(d.device));
    struct SyncOnExit {  // error paths: nothing may still write the caller's arrays
        DevState& d;
        ~SyncOnExit() {
            cudaStreamSynchronize(d.stream);
            cudaGetLastError();
        }
    } sync_on_exit{d};
    uint64_t cap = estimate_capacity(*p, nwin);
    uint64_t abyte_lo = 0;
    size_t anbytes = 0;
    const uint64_t byte_lo = bp_offset / 4 & ~uint64_t(3);
    const size_t nbytes = (size_t)((bp_offset + n_bp + 3) / 4 - byte_lo);
    if ((rc = d.in.reserve(nbytes + 64))) return rc;
    CK(cudaEventRecord(d.ev[0], d.stream));
    CK(cudaMemcpyAsync(d.in.p, packed + byte_lo, nbytes, cudaMemcpyHostToDevice, d.stream));
    if (am.bits && (rc = upload_amb(d, am, 0, n_bp, &abyte_lo, &anbytes))) return rc;
    CK(cudaEventRecord(d.ev[1], d.stream));
    for (int attempt = 0;; attempt++) {  // a too small capacity estimate is re-run with the exact count
        if ((rc = d.pos.reserve(cap))) return rc;
        if (p->want_sk && (rc = d.sk.reserve(cap))) return rc;
        if (vw && (rc = d.val.reserve(cap * vw))) return rc;
        mz::KArgs a{};
        fill_input_args(a, *p, d.in.p, bp_offset, byte_lo, nbytes, nwin);
        if (am.bits) fill_amb_args(a, am, d.amb.p, abyte_lo, anbytes);
        a.pos = d.pos.p, a.sk = d.sk.p, a.val = d.val.p, a.cap = cap;
        if ((rc = enqueue_run(d, *p, a, 0, nwin, &ctx->timing.kernel_launches))) return rc;
        CK(cudaEventRecord(d.ev[2], d.stream));
        CK(cudaStreamSynchronize(d.stream));
        if (!d.hs->overflow) break;
        if (attempt == 1) {
            g_last_error = "internal: exact-capacity re-run overflowed";
            return MZ_ERR_CUDA;
        }
        cap = d.hs->count;
    }
    const uint64_t count = d.hs->count;
    out->count = count;
    if (count > out->capacity) return MZ_ERR_CAPACITY;
    if (count) {
        CK(cudaMemcpyAsync(out->pos, d.pos.p, count * 4, cudaMemcpyDeviceToHost, d.stream));
        if (p->want_sk) CK(cudaMemcpyAsync(out->sk, d.sk.p, count * 4, cudaMemcpyDeviceToHost, d.stream));
        if (vw) CK(cudaMemcpyAsync(out->val, d.val.p, count * 8 * vw, cudaMemcpyDeviceToHost, d.stream));
    }
    CK(cudaEventRecord(d.ev[3], d.stream));
    CK(cudaStreamSynchronize(d.stream));
    cudaEventElapsedTime(&ctx->timing.h2d_ms, d.ev[0], d.ev[1]);
    cudaEventElapsedTime(&ctx->timing.kernel_ms, d.ev[1], d.ev[2]);
    cudaEventElapsedTime(&ctx->timing.d2h_ms, d.ev[2], d.ev[3]);
    cudaEventElapsedTime(&ctx->timing.total_ms, d.ev[0], d.ev[3]);
    ctx->resident.packed = packed, ctx->resident.bp_offset = bp_offset, ctx->resident.n_bp = n_bp;
    ctx->resident.byte_lo = byte_lo, ctx->resident.nbytes = nbytes;
    return MZ_OK;
}

int mz_run(mz_ctx* ctx, const mz_params* p, const uint8_t* packed, uint64_t bp_offset,
           uint64_t n_bp, mz_out* out) {
    return run_host_impl(ctx, p, packed, bp_offset, n_bp, AmbSrc{}, false, out);
}

int mz_run_skip_ambiguous(mz_ctx* ctx, const mz_params* p, const uint8_t* packed, uint64_t bp_offset,
                          uint64_t n_bp, const uint8_t* ambiguous, uint64_t amb_bit_offset, mz_out* out) {
    AmbSrc am;
    am.bits = ambiguous;
    am.off = amb_bit_offset;
    return run_host_impl(ctx, p, packed, bp_offset, n_bp, am, true, out);
}

static int pack_ascii_impl(mz_ctx* ctx, const char* ascii, uint64_t n, uint8_t* packed_out,
                           uint8_t* amb_out, bool want_amb) {
    if (!ctx || (n && (!ascii || !packed_out || (want_amb && !amb_out)))) return MZ_ERR_BAD_ARG;
    if (n == 0) return MZ_OK;
    ctx->resident = mz_ctx::Resident{};
    DevState& d = ctx->devs[0];
    CK(cudaSetDevice(d.device));
    ctx->timing = mz_timing{};
    size_t nbytes = 0, anbytes = 0;
    int rc = pack_on_device(d, ascii, n, &nbytes, &ctx->timing.kernel_launches);
    if (rc) return rc;
    if (want_amb && (rc = amb_on_device(d, n, &anbytes, &ctx->timing.kernel_launches))) return rc;
    CK(cudaMemcpyAsync(packed_out, d.in.p, (n + 3) / 4, cudaMemcpyDeviceToHost, d.stream));
    if (want_amb) CK(cudaMemcpyAsync(amb_out, d.amb.p, (n + 7) / 8, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    return MZ_OK;
}

int mz_pack_ascii(mz_ctx* ctx, const char* ascii, uint64_t n, uint8_t* packed_out) {
    return pack_ascii_impl(ctx, ascii, n, packed_out, nullptr, false);
}

int mz_pack_ascii_n(mz_ctx* ctx, const char* ascii, uint64_t n, uint8_t* packed_out, uint8_t* ambiguous_out) {
    return pack_ascii_impl(ctx, ascii, n, packed_out, ambiguous_out, true);
}

static int run_ascii_impl(mz_ctx* ctx, const mz_params* p, const char* ascii, uint64_t n,
                          bool skip_ambiguous, mz_out* out) {
    if (!ctx || !p || !out) return MZ_ERR_BAD_ARG;
    int rc = mz_params_validate(p, n);
    if (rc) return rc;
    if (skip_ambiguous) {
        if (!p->strand_tiebreak) return MZ_ERR_NOT_CANONICAL;
        if (p->want_sk) return MZ_ERR_BAD_ARG;
    }
    out->count = 0;
    const uint32_t l = p->k + p->w - 1;
    if (n < l) return MZ_OK;
    if (!ascii || !out->pos || (p->want_sk && !out->sk) || (p->value_bits && !out->val)) return MZ_ERR_BAD_ARG;
    ctx->resident = mz_ctx::Resident{};
    DevState& d = ctx->devs[0];
    CK(cudaSetDevice(d.device));
    ctx->timing = mz_timing{};
    size_t nbytes = 0;
    if ((rc = pack_on_device(d, ascii, n, &nbytes, &ctx->timing.kernel_launches))) return rc;
    size_t anbytes = 0;
    if (skip_ambiguous && (rc = amb_on_device(d, n, &anbytes, &ctx->timing.kernel_launches))) return rc;
    const uint64_t nwin = n - l + 1;
    const uint32_t vw = p->value_bits / 64;
    uint64_t cap = estimate_capacity(*p, nwin);
    for (int attempt = 0; attempt < 2; attempt++) {
        if ((rc = d.pos.reserve(cap))) return rc;
        if (p->want_sk && (rc = d.sk.reserve(cap))) return rc;
        if (vw && (rc = d.val.reserve(cap * vw))) return rc;
        mz::KArgs a{};
        fill_input_args(a, *p, d.in.p, 0, 0, nbytes, nwin);
        if (skip_ambiguous) fill_amb_args(a, AmbSrc{}, d.amb.p, 0, anbytes);
        a.pos = d.pos.p, a.sk = d.sk.p, a.val = d.val.p, a.cap = cap;
        if ((rc = enqueue_run(d, *p, a, 0, nwin, &ctx->timing.kernel_launches))) return rc;
        CK(cudaStreamSynchronize(d.stream));
        if (!d.hs->overflow) break;
        if (attempt == 1) {
            g_last_error = "internal: exact-capacity re-run overflowed";
            return MZ_ERR_CUDA;
        }
        cap = d.hs->count;
    }
    const uint64_t count = d.hs->count;
    out->count = count;
    if (count > out->capacity) return MZ_ERR_CAPACITY;
    if (count) {
        CK(cudaMemcpyAsync(out->pos, d.pos.p, count * 4, cudaMemcpyDeviceToHost, d.stream));
        if (p->want_sk) CK(cudaMemcpyAsync(out->sk, d.sk.p, count * 4, cudaMemcpyDeviceToHost, d.stream));
        if (vw) CK(cudaMemcpyAsync(out->val, d.val.p, count * 8 * vw, cudaMemcpyDeviceToHost, d.stream));
    }
    CK(cudaStreamSynchronize(d.stream));
    return MZ_OK;
}

int mz_run_ascii(mz_ctx* ctx, const mz_params* p, const char* ascii, uint64_t n, mz_out* out) {
    return run_ascii_impl(ctx, p, ascii, n, false, out);
}

int mz_run_ascii_skip_ambiguous(mz_ctx* ctx, const mz_params* p, const char* ascii, uint64_t n, mz_out* out) {
    return run_ascii_impl(ctx, p, ascii, n, true, out);
}

int mz_run_batch(mz_ctx* ctx, const mz_params* p, const uint8_t* packed, uint64_t packed_bytes,
                 uint64_t n_reads, const uint64_t* read_start_bp, const uint32_t* read_len_bp,
                 uint64_t stride_bytes, uint32_t fixed_len_bp, uint64_t* out_offsets, mz_out* out) {
    if (!ctx || !p || !out || !out_offsets) return MZ_ERR_BAD_ARG;
    if ((read_start_bp == nullptr) != (read_len_bp == nullptr)) return MZ_ERR_BAD_ARG;
    out->count = 0;
    out_offsets[0] = 0;
    ctx->resident = mz_ctx::Resident{};
    if (n_reads == 0) return MZ_OK;
    const uint32_t l = p->k + p->w - 1;
    uint32_t max_len = fixed_len_bp;
    uint64_t total_windows = 0;
    bool monotone = true;  // ragged reads in storage order: the input can be streamed in chunks
    if (read_len_bp) {
        max_len = 0;
        for (uint64_t r = 0; r < n_reads; r++) {
            max_len = std::max(max_len, read_len_bp[r]);
            total_windows += read_len_bp[r] >= l ? read_len_bp[r] - l + 1 : 0;
            if (2 * (read_start_bp[r] + read_len_bp[r]) > 8 * packed_bytes) return MZ_ERR_BAD_ARG;
            if (r && read_start_bp[r] < read_start_bp[r - 1]) monotone = false;
        }
    } else {
        if (stride_bytes == 0 || (fixed_len_bp + 3) / 4 > stride_bytes) return MZ_ERR_BAD_ARG;
        if ((n_reads - 1) * stride_bytes + (fixed_len_bp + 3) / 4 > packed_bytes) return MZ_ERR_BAD_ARG;
        total_windows = fixed_len_bp >= l ? (uint64_t)(fixed_len_bp - l + 1) * n_reads : 0;
    }
    int rc = mz_params_validate(p, max_len);
    if (rc) return rc;
    if (total_windows == 0) {
        for (uint64_t r = 0; r <= n_reads; r++) out_offsets[r] = 0;
        return MZ_OK;
    }
    if (!packed || !out->pos || (p->want_sk && !out->sk) || (p->value_bits && !out->val)) return MZ_ERR_BAD_ARG;
    const uint32_t S_full = max_len - l + 1;  // windows of the longest read
    const uint32_t vw = p->value_bits / 64;
    CK(cudaSetDevice(ctx->devs[0].device));
    ctx->timing = mz_timing{};
    const auto t_begin = std::chrono::steady_clock::now();
    // Reads are independent units: chunk c of the read list runs on device c % ndev (no halo, no
    // seam), CSR offsets are chunk-local on the device and rebased on the way out.
    const uint64_t ND = ctx->devs.size();
    auto dev_of = [&](uint64_t c) { return (size_t)(c % ND); };
    auto slot_of = [&](uint64_t c) { return (int)((c / ND) % kSlots); };
    std::vector<float> dev_h2d(ND, 0.f), dev_ker(ND, 0.f), dev_d2h(ND, 0.f);
    struct SyncAll {  // error paths: nothing may still write the caller's arrays
        mz_ctx* ctx;
        ~SyncAll() {
            for (size_t i = 0; i < ctx->devs.size(); i++)
                for (int sl = 0; sl < kSlots; sl++) {
                    DevState& d = ctx->slot(i, sl);
                    if (cudaSetDevice(d.device) == cudaSuccess) {
                        cudaStreamSynchronize(d.stream);
                        if (sl == 0) cudaStreamSynchronize(d.up);
                    }
                }
            cudaGetLastError();
            cudaSetDevice(ctx->devs[0].device);
        }
    } sync_all{ctx};

    // Geometry.  A thread handles one read when the longest read fits the per-thread record of
    // the chosen kernel; otherwise reads are cut into pieces of S windows (each piece a thread,
    // seam rule as everywhere: one extra window on the left seeds the dedup comparison).
    const bool lr = p->strand_tiebreak != 0;
    // w > 32: sub-window (XW) instances of the fast kernel
    const bool fast = p->w <= mz::FAST_MAX_W || (p->w <= mz::FAST_XW_MAX_W && !getenv("MZ_NO_XW"));
    uint32_t S_cap, NTg = 128;
    if (fast) {
        // thread = read (or piece of a read): as many windows per thread as the queues hold
        mz::FastPlan probe;
        S_cap = 512;
        while (S_cap > 16 && !mz::fast_queue_plan(S_cap, *p, &probe)) S_cap -= 16;
        if (!mz::fast_queue_plan(S_cap, *p, &probe)) S_cap = 1;
    } else {
        const size_t budget = std::min<size_t>(ctx->devs[0].smem_optin, 200 * 1024);
        while (NTg >= 32 && generic_smem(NTg, 32, p->w, lr) > budget) NTg /= 2;
        if (NTg < 32) {
            g_last_error = "mz_run_batch: w too large for the batch kernels";
            return MZ_ERR_UNSUPPORTED;
        }
        S_cap = 512;
        while (S_cap > 32 && generic_smem(NTg, S_cap, p->w, lr) > budget / 2) S_cap -= 32;
    }
    const uint32_t S = std::min(S_full, S_cap);
    if ((uint64_t)S + p->w + 2 >= 65535) return MZ_ERR_UNSUPPORTED;
    const bool pieces = S_full > S_cap;
    if (pieces && n_reads >= (1ull << 32)) return MZ_ERR_UNSUPPORTED;

    // Reads are processed in chunks that flow through kSlots streams (H2D of chunk c+1 | kernel
    // of c | D2H of c-1 | host copy-out of c-2), like the single-sequence path.  Reads that are
    // not stored in order cannot be streamed: one chunk.
    // ~64 MB of packed input per chunk, smaller (>= 8 MB) when that leaves a device of the context
    // with fewer than four chunks
    const uint64_t chunk_bytes = std::min<uint64_t>(64ull << 20, std::max<uint64_t>(8ull << 20, packed_bytes / (4 * ctx->devs.size())));
    uint64_t chunk_reads = std::max<uint64_t>(1, chunk_bytes / std::max<uint64_t>(1, packed_bytes / n_reads + 1));
    if (const char* e = getenv("MZ_BATCH_CHUNK_READS")) chunk_reads = std::max<uint64_t>(1, strtoull(e, nullptr, 10));
    chunk_reads = (chunk_reads + 15) & ~uint64_t(15);
    if (!monotone || chunk_reads > n_reads) chunk_reads = n_reads;
    const uint64_t nchunks = (n_reads + chunk_reads - 1) / chunk_reads;
    const bool page_in = is_pageable(packed);
    const bool page_out = is_pageable(out->pos) || (p->want_sk && is_pageable(out->sk)) || (vw && is_pageable(out->val));
    const bool page_offs = is_pageable(out_offsets);

    struct BJob {
        uint64_t r0 = 0, r1 = 0, byte_lo = 0, windows = 0, n_units = 0, cap = 0, count = 0, out_off = 0;
        size_t nbytes = 0;
        uint32_t num_tiles = 0, grid = 0;
        size_t smem = 0;
        mz::FastPlan fp;
        mz::KArgs a{};
        bool staged = false;
        std::vector<uint32_t> piece_read, piece_win0;
        const uint8_t* d_in = nullptr;
        cudaEvent_t up_done = nullptr;
    };
    std::vector<BJob> jobs(nchunks);
    uint64_t total = 0;
    bool too_small = false;

    auto launch = [&](DevState& d, BJob& j) -> int {
        int r;
        if ((r = d.pos.reserve(j.cap))) return r;
        if (p->want_sk && (r = d.sk.reserve(j.cap))) return r;
        if (vw && (r = d.val.reserve(j.cap * vw))) return r;
        if ((r = d.scratch.reserve(2 + (size_t)j.num_tiles))) return r;
        CK(cudaMemsetAsync(d.scratch.p, 0, (2 + (size_t)j.num_tiles) * sizeof(unsigned long long), d.stream));
        mz::KArgs& a = j.a;
        a.pos = d.pos.p, a.sk = d.sk.p, a.val = d.val.p, a.cap = j.cap;
        a.count_out = d.scratch.p;
        a.ticket = reinterpret_cast<uint32_t*>(d.scratch.p + 1);
        a.overflow = a.ticket + 1;
        a.tile_state = d.scratch.p + 2;
        attach_host_scalars(d, a);
        if (fast) {
            if ((r = d.rows.reserve((j.fp.scratch_words_per_block * 2 + j.fp.r1_words) * j.fp.grid * mz::FAST_WARPS))) return r;
            a.scratch = d.rows.p;
            mz::fast_plan_args(j.fp, a);
            r = mz::launch_fast(*p, j.fp.grid, a, d.stream);
        } else {
            Plan pl;
            pl.NT = NTg, pl.S = S, pl.smem = j.smem, pl.num_tiles = j.num_tiles, pl.grid = j.grid;
            r = launch_generic(*p, pl, a, d.stream);
        }
        if (r) {
            if (r == MZ_ERR_CUDA && g_last_error.empty()) g_last_error = "kernel launch failed";
            return r;
        }
        ctx->timing.kernel_launches++;
        if (hs_by_copy()) CK(cudaMemcpyAsync(d.hs, d.scratch.p, sizeof(HostScalars), cudaMemcpyDeviceToHost, d.stream));
        return MZ_OK;
    };

    // Fixed-stride reads from pinned memory: the whole input goes up first on a stream of its own
    // (see run_pipelined: copies in both directions at once slow each other down, and the output is
    // 2-3x the input), every chunk's kernel waits for its own piece.
    const bool front = !read_start_bp && !page_in && !getenv("MZ_NO_FRONT_UPLOAD");
    if (front) {
        std::vector<size_t> share(ND, 0), off(ND, 0), nth(ND, 0);
        auto span = [&](uint64_t c, uint64_t* lo, size_t* nb) {
            const uint64_t r0 = c * chunk_reads, r1 = std::min<uint64_t>(r0 + chunk_reads, n_reads);
            *lo = (r0 * stride_bytes) & ~uint64_t(3);
            *nb = (size_t)((r1 - 1) * stride_bytes + (fixed_len_bp + 3) / 4 - *lo);
        };
        for (uint64_t c = 0; c < nchunks; c++) {
            uint64_t lo;
            size_t nb;
            span(c, &lo, &nb);
            share[dev_of(c)] += (nb + 64 + 255) & ~size_t(255);
        }
        for (size_t i = 0; i < ND; i++) {
            CK(cudaSetDevice(ctx->devs[i].device));
            if ((rc = ctx->devs[i].in_all.reserve(share[i] + 256))) return rc;
        }
        for (uint64_t c = 0; c < nchunks; c++) {
            const size_t di = dev_of(c);
            DevState& d0 = ctx->devs[di];
            uint64_t lo;
            size_t nb;
            span(c, &lo, &nb);
            CK(cudaSetDevice(d0.device));
            uint8_t* dst = d0.in_all.p + off[di];
            off[di] += (nb + 64 + 255) & ~size_t(255);
            CK(cudaMemcpyAsync(dst, packed + lo, nb, cudaMemcpyHostToDevice, d0.up));
            if ((rc = upload_event(d0, nth[di]++, &jobs[c].up_done))) return rc;
            CK(cudaEventRecord(jobs[c].up_done, d0.up));
            jobs[c].d_in = dst;
        }
    }

    auto issue = [&](uint64_t c) -> int {
        DevState& d = ctx->slot(dev_of(c), slot_of(c));
        BJob& j = jobs[c];
        int r;
        CK(cudaSetDevice(d.device));
        j.r0 = c * chunk_reads;
        j.r1 = std::min<uint64_t>(j.r0 + chunk_reads, n_reads);
        const uint64_t nr = j.r1 - j.r0;
        // byte range of the chunk's reads (start aligned down to a 32-bit word)
        uint64_t lo, hi;
        if (read_start_bp) {
            if (monotone) {
                lo = read_start_bp[j.r0] / 4;
                hi = lo;
                for (uint64_t r2 = j.r0; r2 < j.r1; r2++) {
                    hi = std::max<uint64_t>(hi, (read_start_bp[r2] + read_len_bp[r2] + 3) / 4);
                    j.windows += read_len_bp[r2] >= l ? read_len_bp[r2] - l + 1 : 0;
                }
            } else {
                lo = 0, hi = packed_bytes, j.windows = total_windows;
            }
        } else {
            lo = j.r0 * stride_bytes;
            hi = (j.r1 - 1) * stride_bytes + (fixed_len_bp + 3) / 4;
            j.windows = fixed_len_bp >= l ? (uint64_t)(fixed_len_bp - l + 1) * nr : 0;
        }
        j.byte_lo = lo & ~uint64_t(3);
        j.nbytes = (size_t)(hi - j.byte_lo);
        j.n_units = nr;
        if (pieces) {
            for (uint64_t r2 = j.r0; r2 < j.r1; r2++) {
                const uint32_t len = read_len_bp ? read_len_bp[r2] : fixed_len_bp;
                const uint32_t nw = len >= l ? len - l + 1 : 0;
                uint32_t w0 = 0;
                do {  // every read owns at least one piece (it writes the read's CSR offset)
                    j.piece_read.push_back((uint32_t)(r2 - j.r0));
                    j.piece_win0.push_back(w0);
                    w0 += S;
                } while (w0 < nw);
            }
            j.n_units = j.piece_read.size();
        }
        if (fast) {
            const uint64_t tiles = (j.n_units + 31) / 32;
            if (tiles > 0x7fffffffull) return MZ_ERR_UNSUPPORTED;
            if (!mz::fast_queue_plan(S, *p, &j.fp)) return MZ_ERR_UNSUPPORTED;
            j.fp.num_tiles = (uint32_t)tiles;
            j.fp.grid = (uint32_t)std::min<uint64_t>(tiles, (uint64_t)d.sm_count * mz::FAST_BPS);
            j.num_tiles = j.fp.num_tiles;
        } else {
            const uint64_t tiles = (j.n_units + NTg - 1) / NTg;
            if (tiles > 0x7fffffffull) return MZ_ERR_UNSUPPORTED;
            j.smem = generic_smem(NTg, S, p->w, lr);
            j.num_tiles = (uint32_t)tiles;
            j.grid = (uint32_t)std::min<uint64_t>(tiles, (uint64_t)d.sm_count * std::max<size_t>(1, std::min<size_t>(8, (220 * 1024) / (j.smem + 1024))));
        }
        j.cap = std::min<uint64_t>(estimate_capacity(*p, j.windows) + nr, std::max<uint64_t>(j.windows, 1));
        if (!front && (r = d.in.reserve(j.nbytes + 64))) return r;
        if ((r = d.offs.reserve(nr + 1))) return r;
        if (read_start_bp) {
            if ((r = d.rstart.reserve(nr))) return r;
            if ((r = d.rlen.reserve(nr))) return r;
        }
        if (pieces) {
            if ((r = d.pread.reserve(j.n_units))) return r;
            if ((r = d.pwin.reserve(j.n_units))) return r;
        }
        const uint8_t* src = packed + j.byte_lo;
        if (page_in && j.nbytes >= (1u << 20)) {
            if ((r = d.st_in.reserve(j.nbytes))) return r;
            parallel_memcpy(d.st_in.p, src, j.nbytes);
            src = d.st_in.p;
        }
        CK(cudaEventRecord(d.ev[0], d.stream));
        if (front) {
            CK(cudaStreamWaitEvent(d.stream, j.up_done, 0));
        } else {
            CK(cudaMemcpyAsync(d.in.p, src, j.nbytes, cudaMemcpyHostToDevice, d.stream));
            j.d_in = d.in.p;
        }
        if (pieces) {
            CK(cudaMemcpyAsync(d.pread.p, j.piece_read.data(), j.n_units * 4, cudaMemcpyHostToDevice, d.stream));
            CK(cudaMemcpyAsync(d.pwin.p, j.piece_win0.data(), j.n_units * 4, cudaMemcpyHostToDevice, d.stream));
        }
        if (read_start_bp) {
            CK(cudaMemcpyAsync(d.rstart.p, read_start_bp + j.r0, nr * 8, cudaMemcpyHostToDevice, d.stream));
            CK(cudaMemcpyAsync(d.rlen.p, read_len_bp + j.r0, nr * 4, cudaMemcpyHostToDevice, d.stream));
        }
        CK(cudaEventRecord(d.ev[1], d.stream));
        mz::KArgs& a = j.a;
        fill_hash_args(a, *p);
        a.seq = reinterpret_cast<const uint32_t*>(j.d_in);
        // ragged reads keep their absolute start positions; fixed-stride reads are chunk-local
        a.bitbias = read_start_bp ? -(int64_t)(8 * j.byte_lo) : (int64_t)(8 * (lo - j.byte_lo));
        a.seq_nwords = (j.nbytes + 3) / 4;
        a.nwin = j.windows;
        a.S = S;
        a.num_tiles = j.num_tiles;
        a.n_reads = j.n_units;
        a.piece_read = pieces ? d.pread.p : nullptr;
        a.piece_win0 = pieces ? d.pwin.p : nullptr;
        a.read_start_bp = read_start_bp ? d.rstart.p : nullptr;
        a.read_len_bp = read_start_bp ? d.rlen.p : nullptr;
        a.stride_bits = stride_bytes * 8;
        a.fixed_len_bp = fixed_len_bp;
        a.out_offsets = d.offs.p;
        if ((r = launch(d, j))) return r;
        CK(cudaEventRecord(d.ev[2], d.stream));
        CK(cudaEventRecord(d.ev[5], d.stream));
        return MZ_OK;
    };
    // kernel done -> counts known -> D2H of this chunk's outputs and CSR offsets
    auto retire = [&](uint64_t c) -> int {
        DevState& d = ctx->slot(dev_of(c), slot_of(c));
        BJob& j = jobs[c];
        int r;
        CK(cudaSetDevice(d.device));
        CK(cudaStreamSynchronize(d.stream));
        float h2d = 0;
        cudaEventElapsedTime(&h2d, d.ev[0], d.ev[1]);
        dev_h2d[dev_of(c)] += h2d;
        dev_ker[dev_of(c)] += chunk_kernel_ms(d, c >= ND ? &ctx->slot(dev_of(c), slot_of(c - ND)) : nullptr);
        if (d.hs->overflow) {  // capacity estimate too small: redo this chunk with the exact size
            j.cap = d.hs->count;
            if ((r = launch(d, j))) return r;
            CK(cudaStreamSynchronize(d.stream));
            if (d.hs->overflow) {
                g_last_error = "internal: exact-capacity re-run overflowed";
                return MZ_ERR_CUDA;
            }
        }
        const uint64_t count = d.hs->count, nr = j.r1 - j.r0;
        if (total + count > out->capacity) too_small = true;
        j.count = count;
        j.out_off = total;
        total += count;
        if (too_small) return MZ_OK;
        CK(cudaEventRecord(d.ev[2], d.stream));
        // CSR offsets: add the chunk's base on the device, then straight into the caller's array
        // (through a pinned bounce buffer when that array is pageable)
        if (j.out_off) {
            const unsigned nt = 256;
            mz_rebase_offsets_kernel<<<(unsigned)((nr + 1 + nt - 1) / nt), nt, 0, d.stream>>>(
                reinterpret_cast<unsigned long long*>(d.offs.p), nr + 1, (unsigned long long)j.out_off);
            CK(cudaGetLastError());
            ctx->timing.kernel_launches++;
        }
        if (page_offs) {
            if ((r = d.st_offs.reserve((nr + 1) * 8))) return r;
            CK(cudaMemcpyAsync(d.st_offs.p, d.offs.p, (nr + 1) * 8, cudaMemcpyDeviceToHost, d.stream));
        } else {
            CK(cudaMemcpyAsync(out_offsets + j.r0 + 1, d.offs.p + 1, nr * 8, cudaMemcpyDeviceToHost, d.stream));
        }
        if (count) {
            uint32_t *hpos = out->pos + j.out_off, *hsk = p->want_sk ? out->sk + j.out_off : nullptr;
            uint64_t* hval = vw ? out->val + j.out_off * vw : nullptr;
            if (page_out) {
                if ((r = d.st_pos.reserve(count * 4))) return r;
                if (p->want_sk && (r = d.st_sk.reserve(count * 4))) return r;
                if (vw && (r = d.st_val.reserve(count * 8 * vw))) return r;
                hpos = reinterpret_cast<uint32_t*>(d.st_pos.p);
                hsk = reinterpret_cast<uint32_t*>(d.st_sk.p);
                hval = reinterpret_cast<uint64_t*>(d.st_val.p);
                j.staged = true;
            }
            CK(cudaMemcpyAsync(hpos, d.pos.p, count * 4, cudaMemcpyDeviceToHost, d.stream));
            if (p->want_sk) CK(cudaMemcpyAsync(hsk, d.sk.p, count * 4, cudaMemcpyDeviceToHost, d.stream));
            if (vw) CK(cudaMemcpyAsync(hval, d.val.p, count * 8 * vw, cudaMemcpyDeviceToHost, d.stream));
        }
        CK(cudaEventRecord(d.ev[3], d.stream));
        return MZ_OK;
    };
    // D2H done -> copy out of the bounce buffers, rebase the CSR offsets
    auto finish = [&](uint64_t c) -> int {
        DevState& d = ctx->slot(dev_of(c), slot_of(c));
        BJob& j = jobs[c];
        if (too_small) return MZ_OK;
        CK(cudaSetDevice(d.device));
        CK(cudaEventSynchronize(d.ev[3]));
        float d2h = 0;
        cudaEventElapsedTime(&d2h, d.ev[2], d.ev[3]);
        dev_d2h[dev_of(c)] += d2h;
        const uint64_t nr = j.r1 - j.r0;
        if (page_offs) parallel_memcpy(out_offsets + j.r0 + 1, reinterpret_cast<const uint64_t*>(d.st_offs.p) + 1, nr * 8);
        if (j.staged && j.count) {
            parallel_memcpy(out->pos + j.out_off, d.st_pos.p, j.count * 4);
            if (p->want_sk) parallel_memcpy(out->sk + j.out_off, d.st_sk.p, j.count * 4);
            if (vw) parallel_memcpy(out->val + j.out_off * vw, d.st_val.p, j.count * 8 * vw);
        }
        std::vector<uint32_t>().swap(j.piece_read);
        std::vector<uint32_t>().swap(j.piece_win0);
        return MZ_OK;
    };
    for (uint64_t c = 0; c < nchunks + 2 * ND; c++) {
        if (c < nchunks && (rc = issue(c))) return rc;
        if (c >= ND && c - ND < nchunks && (rc = retire(c - ND))) return rc;
        if (c >= 2 * ND && (rc = finish(c - 2 * ND))) return rc;
    }
    for (size_t i = 0; i < ND; i++) {  // per-phase times: the busiest device (sums over its chunks)
        ctx->timing.h2d_ms = std::max(ctx->timing.h2d_ms, dev_h2d[i]);
        ctx->timing.kernel_ms = std::max(ctx->timing.kernel_ms, dev_ker[i]);
        ctx->timing.d2h_ms = std::max(ctx->timing.d2h_ms, dev_d2h[i]);
    }
    ctx->timing.total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    out->count = total;
    return too_small ? MZ_ERR_CAPACITY : MZ_OK;
}

}  // extern "C"
