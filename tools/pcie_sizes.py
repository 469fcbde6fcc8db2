import ctypes as C, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sm = importlib.import_module("simd-minimizers_b200"); ffi = importlib.import_module("simd-minimizers_b200._ffi"); L = ffi.lib()
ctx = sm.Context()
for (b, reps) in ((256 << 20, 4), (256 << 20, 16), (1 << 30, 1), (1 << 30, 4), (3 << 30, 1), (3 << 30, 2), (64 << 20, 16)):
    r = ffi.MzPcieResult()
    ffi.check(L.mz_pcie_probe(ctx.handle, b, reps, C.byref(r)))
    print(f"{b >> 20} MiB x {reps}: h2d {r.h2d_gbs:.1f} d2h {r.d2h_gbs:.1f} both {r.bidir_h2d_gbs:.1f} + {r.bidir_d2h_gbs:.1f}")
