// Skip-ambiguous-windows instances of the W-specialised kernel for W = 1 .. 8.
#include "mz_fast.cuh"
namespace mz {
int launch_fast_a0(const mz_params& p, uint32_t grid, const KArgs& a, cudaStream_t st) {
    switch (p.w) {
        case 1: return launch_fast_amb_w<1>(p, grid, a, st);
        case 2: return launch_fast_amb_w<2>(p, grid, a, st);
        case 3: return launch_fast_amb_w<3>(p, grid, a, st);
        case 4: return launch_fast_amb_w<4>(p, grid, a, st);
        case 5: return launch_fast_amb_w<5>(p, grid, a, st);
        case 6: return launch_fast_amb_w<6>(p, grid, a, st);
        case 7: return launch_fast_amb_w<7>(p, grid, a, st);
        case 8: return launch_fast_amb_w<8>(p, grid, a, st);
        default: return MZ_ERR_UNSUPPORTED;
    }
}
}  // namespace mz
