import ctypes as C, importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
sm = importlib.import_module("simd-minimizers_b200"); ffi = importlib.import_module("simd-minimizers_b200._ffi"); L = ffi.lib()
cfg = bench.CONFIGS["c2"]; n = cfg["n"]
host, off = bench.synth_packed_range(bench.SEED, 0, n)
pin = bench.HostBuf(L, ffi, host.size, np.uint8); pin.np[:host.size] = host; del host
k, w = cfg["k"], cfg["w"]
p = ffi.MzParams(); L.mz_params_nthash(C.byref(p), k, w, 0, 1); p.value_bits = 64
cap = int(n * 0.1 * 1.1) + 65536
h_pos = bench.HostBuf(L, ffi, cap * 4, np.uint32); h_val = bench.HostBuf(L, ffi, cap * 8, np.uint64)
ctx = sm.Context([0])
for env in ({}, {"MZ_HOST_THREADS": "4"}, {"MZ_NO_FRONT_UPLOAD": "1"}, {"MZ_CHUNK_WINDOWS": "64000000"}, {"MZ_CHUNK_WINDOWS": "260000000"}):
    for kk in ("MZ_HOST_THREADS", "MZ_NO_FRONT_UPLOAD", "MZ_CHUNK_WINDOWS"): os.environ.pop(kk, None)
    os.environ.update(env)
    print("==", env, flush=True)
    for it in range(3):
        out = ffi.MzOut(h_pos.ptr, None, h_val.ptr, cap, 0)
        if it == 2: os.environ["MZ_DEBUG_PIPE"] = "2"
        t0 = time.perf_counter()
        ffi.check(L.mz_run(ctx.handle, C.byref(p), pin.ptr, off, n, C.byref(out)))
        dt = (time.perf_counter() - t0) * 1e3
        os.environ.pop("MZ_DEBUG_PIPE", None)
    print(f"   {dt:.1f} ms", flush=True)
