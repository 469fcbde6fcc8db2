"""Device-resident time of the headline config (canonical k=31 w=19, pos + u64 values) over a range of
input sizes: min of 10 launches each, fit t = a + b n.  MZ_B200_LIB selects another build for A/B."""
import ctypes as C, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
sm = importlib.import_module("simd-minimizers_b200"); ffi = importlib.import_module("simd-minimizers_b200._ffi"); L = ffi.lib()
sizes = [10_000_000, 50_000_000, 100_000_000, 200_000_000, 387_500_000, 800_000_000, 1_600_000_000, 3_100_000_000]
nmax = max(sizes)
host, off = bench.synth_packed_range(bench.SEED, 0, nmax)
d_in = torch.from_numpy(np.ascontiguousarray(host)).cuda(); del host
ctx = sm.Context()
k, w = 31, 19
p = ffi.MzParams(); L.mz_params_nthash(C.byref(p), k, w, 0, 1); p.value_bits = 64
cap = int(nmax * 2.3 / (w + 1)) + 65536
dp = torch.empty(cap, dtype=torch.int32, device="cuda"); dv = torch.empty(cap, dtype=torch.int64, device="cuda")
xs, ys = [], []
for n in sizes:
    ts = []
    for it in range(12):
        out = ffi.MzOut(dp.data_ptr(), None, dv.data_ptr(), cap, 0)
        assert L.mz_run_device(ctx.handle, 0, C.byref(p), d_in.data_ptr(), off, n, 0, 0, C.byref(out)) == 0
        ts.append(ctx.last_timing()["kernel_ms"])
    t = min(ts[2:])
    xs.append(n); ys.append(t)
    print(f"n={n:>11}: {t:8.4f} ms  {n / t / 1e6:7.1f} Gbp/s  (median {sorted(ts[2:])[5]:.4f})", flush=True)
b, a = np.polyfit(np.array(xs[3:], dtype=float), np.array(ys[3:]), 1)
print(f"fit over n >= {xs[3]}: t = {a * 1e3:.1f} us + n / ({1e-6 / b:.1f} Gbp/s)")
