"""min-of-N device-resident time for the headline config on 800 Mbp (A/B testing helper)."""
import ctypes as C, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
sm = importlib.import_module("simd-minimizers_b200"); ffi = importlib.import_module("simd-minimizers_b200._ffi"); L = ffi.lib()
n = 800_000_000
host, off = bench.synth_packed_range(bench.SEED, 0, n)
d_in = torch.from_numpy(np.ascontiguousarray(host)).cuda()
ctx = sm.Context()
for (k, w, canon, vb) in ((31, 19, 1, 64), (21, 11, 0, 0)):
    p = ffi.MzParams(); L.mz_params_nthash(C.byref(p), k, w, 0, canon); p.value_bits = vb
    cap = int(n * 2.4 / (w + 1)) + 65536
    dp = torch.empty(cap, dtype=torch.int32, device="cuda"); dv = torch.empty(cap if vb else 1, dtype=torch.int64, device="cuda")
    ts = []
    for it in range(12):
        out = ffi.MzOut(dp.data_ptr(), None, dv.data_ptr() if vb else None, cap, 0)
        assert L.mz_run_device(ctx.handle, 0, C.byref(p), d_in.data_ptr(), off, n, 0, 0, C.byref(out)) == 0
        ts.append(ctx.last_timing()["kernel_ms"])
    ts = sorted(ts[2:])
    print(f"k={k} w={w} canon={canon} vals={vb}: min {ts[0]:.3f} ms ({n/ts[0]/1e6:.1f} Gbp/s)  median {ts[len(ts)//2]:.3f} ms  max {ts[-1]:.3f}")
