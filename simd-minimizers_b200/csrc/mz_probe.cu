// mz_probe.cu -- INT32 ALU-pipe micro-benchmark: the measured
// denominator of the integer roofline bench.py reports next to the HBM roofline.
//
// The minimizer kernels are bound by the integer ALU pipe (LOP3 / SHF / VIMNMX / IADD3 / ISETP /
// PRMT all issue there); its peak is not in MEASURED_PEAKS.json, so it is measured here, on the
// device and at the clocks the bench runs at: every thread runs eight independent dependency
// chains of {lop3, shf.l.wrap, min.u32, prmt} (the ALU-pipe mix of the van-Herk loop; plain adds are
// left out because ptxas moves them to the FMA pipe as IMAD.IADD whenever the ALU pipe is the busy one), enough warps per SM
// to hide the 4-cycle ALU latency, no memory traffic.
#include "../../include/mz_b200.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int kChains = 8;
constexpr int kOpsPerIter = 4 * kChains;

__global__ void __launch_bounds__(1024, 2) mz_alu_probe_kernel(uint32_t iters, uint32_t seed, uint32_t* sink) {
    uint32_t a[kChains], b[kChains];
#pragma unroll
    for (int c = 0; c < kChains; c++) {
        a[c] = seed * (2u * c + 1u) + threadIdx.x;
        b[c] = (seed ^ 0x9e3779b9u) + blockIdx.x * 977u + c;
    }
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < kChains; c++) {
            uint32_t t, u, v, x;
            asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(t) : "r"(a[c]), "r"(b[c]), "r"(seed));  // LOP3
            asm volatile("shf.l.wrap.b32 %0, %1, %1, 7;" : "=r"(u) : "r"(t));                           // SHF
            asm volatile("min.u32 %0, %1, %2;" : "=r"(v) : "r"(u), "r"(b[c]));                          // VIMNMX
            asm volatile("prmt.b32 %0, %1, %2, 0x2165;" : "=r"(x) : "r"(v), "r"(t));                    // PRMT
            a[c] = x;
            b[c] = u;
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int c = 0; c < kChains; c++) r ^= a[c] ^ b[c];
    if (r == 0x12345678u) *sink = r;  // practically never: keeps the chains alive
}

}  // namespace

extern "C" int mz_alu_probe_device(int device, mz_alu_result* res) {
    if (!res) return MZ_ERR_BAD_ARG;
    res->lane_ops_per_s = 0;
    res->ms = 0;
    if (cudaSetDevice(device) != cudaSuccess) return MZ_ERR_NO_DEVICE;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    uint32_t* sink = nullptr;
    if (cudaMalloc(&sink, 4) != cudaSuccess) return MZ_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const uint32_t iters = 4096;
    const unsigned grid = (unsigned)sms * 2, block = 1024;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {  // first launch warms up
        cudaEventRecord(e0);
        mz_alu_probe_kernel<<<grid, block>>>(iters, 12345u + rep, sink);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const cudaError_t e = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    if (e != cudaSuccess || best > 1e29f) return MZ_ERR_CUDA;
    res->ms = best;
    res->lane_ops_per_s = (double)grid * block * (double)iters * kOpsPerIter / (best * 1e-3);
    res->sm_count = (uint32_t)sms;
    return MZ_OK;
}
