"""e2e time of mz_run with pageable (ordinary malloc) host buffers vs pinned ones."""
import ctypes as C, importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
sm = importlib.import_module("simd-minimizers_b200"); ffi = importlib.import_module("simd-minimizers_b200._ffi"); L = ffi.lib()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_100_000_000
host, off = bench.synth_packed_range(bench.SEED, 0, n)
host = np.ascontiguousarray(host)
k, w = 31, 19
p = ffi.MzParams(); L.mz_params_nthash(C.byref(p), k, w, 0, 1); p.value_bits = 64
cap = int(n * 2.3 / (w + 1)) + 65536
ctx = sm.Context()
def run(inp_ptr, pos_ptr, val_ptr, label):
    ts = []
    for it in range(4):
        out = ffi.MzOut(pos_ptr, None, val_ptr, cap, 0)
        t0 = time.perf_counter()
        rc = L.mz_run(ctx.handle, C.byref(p), inp_ptr, off, n, C.byref(out)); assert rc == 0, rc
        ts.append(time.perf_counter() - t0)
    print(f"{label}: best {min(ts[1:])*1e3:.1f} ms = {n/min(ts[1:])/1e9:.1f} Gbp/s (first call {ts[0]*1e3:.0f} ms), count {out.count}")
pos = np.empty(cap, dtype=np.uint32); val = np.empty(cap, dtype=np.uint64)
pos[:] = 0; val[:] = 0
run(host.ctypes.data, pos.ctypes.data, val.ctypes.data, "pageable in/out")
hin = torch.from_numpy(host).pin_memory(); hpos = torch.empty(cap, dtype=torch.int32).pin_memory(); hval = torch.empty(cap, dtype=torch.int64).pin_memory()
run(hin.data_ptr(), hpos.data_ptr(), hval.data_ptr(), "pinned in/out")
assert np.array_equal(hpos.numpy().view(np.uint32)[:1000], pos[:1000])
