// Long-window (XW) instances of the W-specialised kernel, sub-window length 21 .. 24.
#include "mz_fast.cuh"
namespace mz {
int launch_fast_x1(const mz_params& p, uint32_t grid, const KArgs& a, cudaStream_t st) {
    switch (fast_wt(p.w)) {
        case 21: return launch_fast_xw_w<21>(p, grid, a, st);
        case 22: return launch_fast_xw_w<22>(p, grid, a, st);
        case 23: return launch_fast_xw_w<23>(p, grid, a, st);
        case 24: return launch_fast_xw_w<24>(p, grid, a, st);
        default: return MZ_ERR_UNSUPPORTED;
    }
}
}  // namespace mz
