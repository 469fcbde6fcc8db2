"""Device-resident timing of the generic (runtime-w) kernel: w > 255."""
import ctypes as C, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
sm = importlib.import_module("simd-minimizers_b200"); ffi = importlib.import_module("simd-minimizers_b200._ffi"); L = ffi.lib()
n = 100_000_000
host, off = bench.synth_packed_range(bench.SEED, 0, n)
d_in = torch.from_numpy(np.ascontiguousarray(host)).cuda()
ctx = sm.Context()
for (k, w, canon, vb) in ((21, 301, 1, 64), (21, 1001, 1, 64), (15, 4000, 0, 0)):
    p = ffi.MzParams(); L.mz_params_nthash(C.byref(p), k, w, 0, canon); p.value_bits = vb
    cap = int(n * 2.4 / (w + 1)) + 65536
    dp = torch.empty(cap, dtype=torch.int32, device="cuda"); dv = torch.empty(cap if vb else 1, dtype=torch.int64, device="cuda")
    ts = []
    for it in range(3):
        out = ffi.MzOut(dp.data_ptr(), None, dv.data_ptr() if vb else None, cap, 0)
        assert L.mz_run_device(ctx.handle, 0, C.byref(p), d_in.data_ptr(), off, n, 0, 0, C.byref(out)) == 0
        ts.append(ctx.last_timing()["kernel_ms"])
    t = min(ts[1:])
    print(f"S={os.environ.get('MZ_GENERIC_S','auto')} gring={os.environ.get('MZ_GENERIC_GRING','auto')} k={k} w={w} canon={canon}: {t:.3f} ms ({n/t/1e6:.1f} Gbp/s)")
