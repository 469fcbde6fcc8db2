// Skip-ambiguous-windows instances of the W-specialised kernel for W = 17 .. 24.
#include "mz_fast.cuh"
namespace mz {
int launch_fast_a2(const mz_params& p, uint32_t grid, const KArgs& a, cudaStream_t st) {
    switch (p.w) {
        case 17: return launch_fast_amb_w<17>(p, grid, a, st);
        case 18: return launch_fast_amb_w<18>(p, grid, a, st);
        case 19: return launch_fast_amb_w<19>(p, grid, a, st);
        case 20: return launch_fast_amb_w<20>(p, grid, a, st);
        case 21: return launch_fast_amb_w<21>(p, grid, a, st);
        case 22: return launch_fast_amb_w<22>(p, grid, a, st);
        case 23: return launch_fast_amb_w<23>(p, grid, a, st);
        case 24: return launch_fast_amb_w<24>(p, grid, a, st);
        default: return MZ_ERR_UNSUPPORTED;
    }
}
}  // namespace mz
