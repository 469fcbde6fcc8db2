"""Device-resident throughput over a (k, w, builder) matrix on 1 Gbp (perf-cliff check)."""
import ctypes as C, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
sm = importlib.import_module("simd-minimizers_b200"); ffi = importlib.import_module("simd-minimizers_b200._ffi"); L = ffi.lib()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
host, off = bench.synth_packed_range(bench.SEED, 0, n)
d_in = torch.from_numpy(np.ascontiguousarray(host)).cuda()
ctx = sm.Context()
cases = [(21, 11, 0, 0, 0), (21, 11, 1, 0, 64), (31, 19, 1, 0, 64), (31, 5, 1, 0, 64), (15, 10, 0, 0, 0), (19, 19, 1, 0, 0), (11, 31, 1, 0, 64),
         (31, 32, 0, 0, 0), (8, 3, 1, 0, 64), (5, 1, 1, 0, 64), (31, 11, 1, 1, 0), (31, 11, 1, 2, 0), (21, 41, 1, 0, 64), (31, 19, 1, 0, 0),
         (15, 50, 0, 0, 0), (31, 63, 1, 0, 64), (21, 101, 1, 0, 64), (21, 201, 1, 0, 64), (21, 301, 1, 0, 64)]
for (k, w, canon, mode, vb) in cases:
    if canon and (k + w - 1) % 2 == 0: w += 1
    p = ffi.MzParams(); L.mz_params_nthash(C.byref(p), k, w, mode, canon); p.value_bits = vb
    nwin = n - (k + w - 1) + 1
    dens = 2.0 / (w + 1) if mode == 0 else (min(1.0, 2.0 / w) if mode == 1 else 1.0 / w)
    cap = int(nwin * min(1.0, dens * 1.2)) + 65536
    dp = torch.empty(cap, dtype=torch.int32, device="cuda"); dv = torch.empty(cap if vb else 1, dtype=torch.int64, device="cuda")
    best = 1e9
    for it in range(4):
        out = ffi.MzOut(dp.data_ptr(), None, dv.data_ptr() if vb else None, cap, 0)
        rc = L.mz_run_device(ctx.handle, 0, C.byref(p), d_in.data_ptr(), off, n, 0, 0, C.byref(out))
        assert rc == 0, rc
        best = min(best, ctx.last_timing()["kernel_ms"])
    print(f"k={k:2d} w={w:2d} canon={canon} mode={mode} vals={vb:3d}: {best:7.3f} ms  {n/best/1e6:7.1f} Gbp/s  density {out.count/nwin:.4f}")
