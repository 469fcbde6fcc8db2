// Long-window (XW) instances of the W-specialised kernel, sub-window length 17 .. 20.
#include "mz_fast.cuh"
namespace mz {
int launch_fast_x0(const mz_params& p, uint32_t grid, const KArgs& a, cudaStream_t st) {
    switch (fast_wt(p.w)) {
        case 17: return launch_fast_xw_w<17>(p, grid, a, st);
        case 18: return launch_fast_xw_w<18>(p, grid, a, st);
        case 19: return launch_fast_xw_w<19>(p, grid, a, st);
        case 20: return launch_fast_xw_w<20>(p, grid, a, st);
        default: return MZ_ERR_UNSUPPORTED;
    }
}
}  // namespace mz
