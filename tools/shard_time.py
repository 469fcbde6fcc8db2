"""Device-resident time of the headline config on the shard sizes of a strong-scaling run (3.1 Gbp / N)."""
import ctypes as C, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
sm = importlib.import_module("simd-minimizers_b200"); ffi = importlib.import_module("simd-minimizers_b200._ffi"); L = ffi.lib()
ctx = sm.Context()
k, w = 31, 19
for N in (8, 4, 2):
    n = 3_100_000_000 // N
    host, off = bench.synth_packed_range(bench.SEED, 0, n)
    d_in = torch.from_numpy(np.ascontiguousarray(host)).cuda()
    p = ffi.MzParams(); L.mz_params_nthash(C.byref(p), k, w, 0, 1); p.value_bits = 64
    cap = int(n * 2.4 / (w + 1)) + 65536
    dp = torch.empty(cap, dtype=torch.int32, device="cuda"); dv = torch.empty(cap, dtype=torch.int64, device="cuda")
    ts = []
    for it in range(12):
        out = ffi.MzOut(dp.data_ptr(), None, dv.data_ptr(), cap, 0)
        assert L.mz_run_device(ctx.handle, 0, C.byref(p), d_in.data_ptr(), off, n, 0, 0, C.byref(out)) == 0
        ts.append(ctx.last_timing()["kernel_ms"])
    ts = sorted(ts[2:])
    print(f"3.1 Gbp / {N} = {n} bases: min {ts[0]:.4f} ms ({n/ts[0]/1e6:.1f} Gbp/s, x{N} = {N*n/ts[0]/1e6:.0f})  median {ts[len(ts)//2]:.4f} ms")
    del d_in, dp, dv
