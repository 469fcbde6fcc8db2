"""One device-resident launch configuration of the long-window path, repeated (for ncu)."""
import ctypes as C, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
sm = importlib.import_module("simd-minimizers_b200"); ffi = importlib.import_module("simd-minimizers_b200._ffi"); L = ffi.lib()
k, w, canon, vb = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (21, 41, 1, 64)))
n = 400_000_000
host, off = bench.synth_packed_range(bench.SEED, 0, n)
d_in = torch.from_numpy(np.ascontiguousarray(host)).cuda()
ctx = sm.Context()
p = ffi.MzParams(); L.mz_params_nthash(C.byref(p), k, w, 0, canon); p.value_bits = vb
cap = int(n * 2.4 / (w + 1)) + 65536
dp = torch.empty(cap, dtype=torch.int32, device="cuda"); dv = torch.empty(cap if vb else 1, dtype=torch.int64, device="cuda")
for it in range(4):
    out = ffi.MzOut(dp.data_ptr(), None, dv.data_ptr() if vb else None, cap, 0)
    assert L.mz_run_device(ctx.handle, 0, C.byref(p), d_in.data_ptr(), off, n, 0, 0, C.byref(out)) == 0
    print(ctx.last_timing()["kernel_ms"])
