"""Device-resident throughput of the skip-ambiguous kernel instances (800 Mbp, canonical k=31 w=19,
pos + u64 values) for three masks: empty, genome-like N runs, 1 % isolated N."""
import ctypes as C, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
sm = importlib.import_module("simd-minimizers_b200"); ffi = importlib.import_module("simd-minimizers_b200._ffi"); L = ffi.lib()
n = 800_000_000
host, off = bench.synth_packed_range(bench.SEED, 0, n)
d_in = torch.from_numpy(np.ascontiguousarray(host)).cuda()
ctx = sm.Context()
g = torch.Generator(device="cuda"); g.manual_seed(1)
masks = {}
masks["empty"] = torch.zeros(n // 8 + 64, dtype=torch.uint8, device="cuda")
m = torch.zeros(n // 8 + 64, dtype=torch.uint8, device="cuda")
starts = torch.randint(0, n // 8 - 200_000, (60,), generator=g, device="cuda").tolist()
lens = torch.randint(1, 150_000, (60,), generator=g, device="cuda").tolist()
for s, ln in zip(starts, lens):
    m[s:s + ln] = 0xff          # ~60 runs of up to 1.2 Mbp of N (about 4.5 % of the sequence)
masks["runs"] = m
r = torch.randint(0, 100, (n // 8 + 64,), generator=g, device="cuda")
bit = torch.randint(0, 8, (n // 8 + 64,), generator=g, device="cuda")
masks["1pct"] = torch.where(r < 8, (1 << bit), 0).to(torch.uint8)   # ~1 % of bases
k, w = 31, 19
p = ffi.MzParams(); L.mz_params_nthash(C.byref(p), k, w, 0, 1); p.value_bits = 64
cap = int(n * 2.4 / (w + 1)) + 65536
dp = torch.empty(cap, dtype=torch.int32, device="cuda"); dv = torch.empty(cap, dtype=torch.int64, device="cuda")
def timeit(fn):
    ts = []
    for it in range(10):
        out = ffi.MzOut(dp.data_ptr(), None, dv.data_ptr(), cap, 0)
        assert fn(out) == 0
        ts.append(ctx.last_timing()["kernel_ms"])
    return sorted(ts[2:])[0], out.count
t, c = timeit(lambda out: L.mz_run_device(ctx.handle, 0, C.byref(p), d_in.data_ptr(), off, n, 0, 0, C.byref(out)))
print(f"plain run:            {t:.3f} ms  {n/t/1e6:.1f} Gbp/s  count {c}")
for name, mk in masks.items():
    t, c = timeit(lambda out: L.mz_run_device_skip_ambiguous(ctx.handle, 0, C.byref(p), d_in.data_ptr(), off, n,
                                                             mk.data_ptr(), 0, 0, 0, C.byref(out)))
    print(f"skip-ambiguous {name:6s}: {t:.3f} ms  {n/t/1e6:.1f} Gbp/s  count {c}")
