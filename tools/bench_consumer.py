"""End-to-end time of the on-device consumer (mz_run_bucket_stats) next to moving the same run's
super-k-mer starts + values to the host (mz_run), C2-sized input, pinned host memory.
python tools/bench_consumer.py [n_bases] [n_buckets]"""
import ctypes as C, importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
sm = importlib.import_module("simd-minimizers_b200"); ffi = importlib.import_module("simd-minimizers_b200._ffi"); L = ffi.lib()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_100_000_000
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
k, w = 31, 19
ctx = sm.Context()
nbytes = (n + 3) // 4 + 64
hp = C.c_void_p(); assert L.mz_host_alloc(C.byref(hp), nbytes) == 0
host = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(nbytes,))
bench.synth_fill(host[: nbytes // 8 * 8].view(np.uint64), bench.SEED, 0)
p = ffi.MzParams(); L.mz_params_nthash(C.byref(p), k, w, 0, 1)
# consumer
sk_h = np.zeros(nb, np.uint64); win_h = np.zeros(nb, np.uint64); nm = C.c_uint64()
ts = []
for it in range(5):
    t0 = time.perf_counter()
    rc = L.mz_run_bucket_stats(ctx.handle, C.byref(p), hp, 0, n, nb, sk_h.ctypes.data, win_h.ctypes.data, C.byref(nm))
    assert rc == 0, rc
    ts.append(time.perf_counter() - t0)
t_dev = min(ts[1:])
assert int(win_h.sum()) == n - (k + w - 1) + 1 and int(sk_h.sum()) == nm.value
# the same information through the host: sk + values over PCIe (the histogram itself not even counted)
cap = int(n * 2.4 / (w + 1)) + 65536
bufs = []
for sz in (4, 4, 8):
    q = C.c_void_p(); assert L.mz_host_alloc(C.byref(q), cap * sz) == 0; bufs.append(q)
p.want_sk = 1; p.value_bits = 64
ts = []
for it in range(4):
    out = ffi.MzOut(bufs[0].value, bufs[1].value, bufs[2].value, cap, 0)
    t0 = time.perf_counter()
    rc = L.mz_run(ctx.handle, C.byref(p), hp, 0, n, C.byref(out))
    assert rc == 0, rc
    ts.append(time.perf_counter() - t0)
t_host = min(ts[1:])
assert out.count == nm.value
print(f"n = {n} bases, canonical k={k} w={w}, {nm.value} minimizers, {nb} buckets")
print(f"mz_run_bucket_stats: {t_dev*1e3:8.2f} ms = {n/t_dev/1e9:6.1f} Gbp/s end to end; over PCIe: {nbytes/1e6:.0f} MB in, {16*nb/1e3:.0f} KB out")
print(f"mz_run (pos+sk+values to the host): {t_host*1e3:8.2f} ms = {n/t_host/1e9:6.1f} Gbp/s; over PCIe: {nbytes/1e6:.0f} MB in, {out.count*16/1e6:.0f} MB out (before any host-side bucketing)")
