// Skip-ambiguous-windows instances of the W-specialised kernel for W = 25 .. 32.
#include "mz_fast.cuh"
namespace mz {
int launch_fast_a3(const mz_params& p, uint32_t grid, const KArgs& a, cudaStream_t st) {
    switch (p.w) {
        case 25: return launch_fast_amb_w<25>(p, grid, a, st);
        case 26: return launch_fast_amb_w<26>(p, grid, a, st);
        case 27: return launch_fast_amb_w<27>(p, grid, a, st);
        case 28: return launch_fast_amb_w<28>(p, grid, a, st);
        case 29: return launch_fast_amb_w<29>(p, grid, a, st);
        case 30: return launch_fast_amb_w<30>(p, grid, a, st);
        case 31: return launch_fast_amb_w<31>(p, grid, a, st);
        case 32: return launch_fast_amb_w<32>(p, grid, a, st);
        default: return MZ_ERR_UNSUPPORTED;
    }
}
}  // namespace mz
