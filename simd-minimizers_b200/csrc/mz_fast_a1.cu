// Skip-ambiguous-windows instances of the W-specialised kernel for W = 9 .. 16.
#include "mz_fast.cuh"
namespace mz {
int launch_fast_a1(const mz_params& p, uint32_t grid, const KArgs& a, cudaStream_t st) {
    switch (p.w) {
        case 9: return launch_fast_amb_w<9>(p, grid, a, st);
        case 10: return launch_fast_amb_w<10>(p, grid, a, st);
        case 11: return launch_fast_amb_w<11>(p, grid, a, st);
        case 12: return launch_fast_amb_w<12>(p, grid, a, st);
        case 13: return launch_fast_amb_w<13>(p, grid, a, st);
        case 14: return launch_fast_amb_w<14>(p, grid, a, st);
        case 15: return launch_fast_amb_w<15>(p, grid, a, st);
        case 16: return launch_fast_amb_w<16>(p, grid, a, st);
        default: return MZ_ERR_UNSUPPORTED;
    }
}
}  // namespace mz
