"""End-to-end (host pinned in -> host pinned out) time of mz_run through ONE context over 1..N devices,
next to the measured copy ceiling (mz_pcie_probe).  Usage: python tools/e2e_probe.py [config] [n_bases]
Environment knobs are the library's own (MZ_DEBUG_PIPE=1 prints the per-device phase breakdown)."""
import ctypes as C, importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
sm = importlib.import_module("simd-minimizers_b200"); ffi = importlib.import_module("simd-minimizers_b200._ffi"); L = ffi.lib()
cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
n = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["n"]
ndev = torch.cuda.device_count()
host, off = bench.synth_packed_range(bench.SEED, 0, n)
pin = bench.HostBuf(L, ffi, host.size, np.uint8); pin.np[:host.size] = host; del host
k, w = cfg["k"], cfg["w"]
p = ffi.MzParams()
(L.mz_params_mulhash if cfg["hasher"] == "mul" else L.mz_params_nthash)(C.byref(p), k, w, cfg["mode"], int(cfg["canonical"]))
p.want_sk, p.value_bits = cfg["want_sk"], cfg["value_bits"]
vw = cfg["value_bits"] // 64
dens = 2.0 / (w + 1) if cfg["mode"] == 0 else 2.0 / w
cap = int(n * dens * 1.1) + 65536
h_pos = bench.HostBuf(L, ffi, cap * 4, np.uint32)
h_sk = bench.HostBuf(L, ffi, (cap if cfg["want_sk"] else 1) * 4, np.uint32)
h_val = bench.HostBuf(L, ffi, max(cap * vw, 1) * 8, np.uint64)
sets = [list(range(m)) for m in (1, 2, 4, 8) if m <= ndev]
ref = None
for devs in sets:
    ctx = sm.Context(devs)
    r = ffi.MzPcieResult()
    ffi.check(L.mz_pcie_probe(ctx.handle, 256 << 20, 4, C.byref(r)))
    print(f"devices {devs}: pcie h2d {r.h2d_gbs:.1f} d2h {r.d2h_gbs:.1f} GB/s alone; both at once h2d {r.bidir_h2d_gbs:.1f} + d2h {r.bidir_d2h_gbs:.1f} GB/s", flush=True)
    for env in ({}, {"MZ_NO_POS_DELTA": "1"}, {"MZ_DELTA_MAX_DEVICES": "64"}):
        for kk in ("MZ_NO_POS_DELTA", "MZ_DELTA_MAX_DEVICES"):
            os.environ.pop(kk, None)
        os.environ.update(env)
        ts = []
        for it in range(4):
            out = ffi.MzOut(h_pos.ptr, h_sk.ptr if cfg["want_sk"] else None, h_val.ptr if vw else None, cap, 0)
            t0 = time.perf_counter()
            ffi.check(L.mz_run(ctx.handle, C.byref(p), pin.ptr, off, n, C.byref(out)))
            ts.append((time.perf_counter() - t0) * 1e3)
        cs = (int(out.count), int(h_pos.np[:out.count].astype(np.uint64).sum()), int(h_val.np[:out.count * vw].sum()) if vw else 0)
        if ref is None:
            ref = cs
        assert cs == ref, (cs, ref)
        t = ctx.last_timing()
        bytes_out = out.count * (4 + 4 * cfg["want_sk"] + 8 * vw)
        print(f"  {env or 'default'}: {min(ts[1:]):.1f} ms best, {np.median(ts[1:]):.1f} median -> {n / min(ts[1:]) / 1e6:.1f} Gbp/s; "
              f"delivered {bytes_out / 1e9:.2f} GB out + {n / 4e9:.2f} GB in = {(bytes_out + n / 4) / min(ts[1:]) / 1e6:.1f} GB/s; "
              f"busiest device: h2d {t['h2d_ms']:.1f} kernels {t['kernel_ms']:.1f} d2h {t['d2h_ms']:.1f} ms; checksum ok", flush=True)
    ctx.close()
