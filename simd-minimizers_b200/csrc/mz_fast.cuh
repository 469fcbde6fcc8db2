// mz_fast.cuh -- W-specialised register-resident kernel (placeholder: not built yet).
#pragma once
#include "../../include/mz_b200.h"
#include "mz_common.cuh"
namespace mz {
inline bool plan_fast(int, size_t, const mz_params&, uint64_t, uint32_t*, uint32_t*, size_t*, uint32_t*) { return false; }
inline int launch_fast(const mz_params&, uint32_t, size_t, uint32_t, const KArgs&, cudaStream_t) { return MZ_ERR_UNSUPPORTED; }
}  // namespace mz
