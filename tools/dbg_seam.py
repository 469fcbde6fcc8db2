"""Debugging aid: where do the host (chunk-pipelined) path and the device-resident path differ from
the oracle?  Prints the differing positions relative to chunk / tile boundaries."""
import os, sys, importlib, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import mzoracle as o
sm = importlib.import_module("simd-minimizers_b200")
ffi = importlib.import_module("simd-minimizers_b200._ffi"); L = ffi.lib()
import torch
n = int(os.environ.get("DBG_N", 80_000_000))
k, w = 31, 19
packed = o.synth_packed(9, n + 8)
pr = o.make_params(k, w, canonical=True)
b = sm.canonical_minimizers(k, w)
want, _ = o.run(packed, 0, n, pr, "stream")
nwin = n - (k + w - 1) + 1
chunk = int(os.environ.get("MZ_CHUNK_WINDOWS", max(1 << 25, (nwin + 23) // 24)))
print("n", n, "nwin", nwin, "chunk", chunk, "env", {e: os.environ[e] for e in os.environ if e.startswith("MZ_")})
got = b.run_once(sm.PackedSeq(packed, 0, n))
print("host path", len(want), len(got), np.array_equal(want, got))
if not np.array_equal(want, got):
    sw, sg = set(want.tolist()), set(got.tolist())
    miss, extra = sorted(sw - sg), sorted(sg - sw)
    print("  missing", miss[:20], "extra", extra[:20])
    for p in miss[:20]:
        i = int(np.searchsorted(want, p))
        print("   missing pos", p, "entry", i, "pos - chunk*c", p % chunk, "chunk", p // chunk, "neighbours", want[max(0, i - 2):i + 3].tolist())
    # first index where the arrays differ
    m = min(len(want), len(got))
    d = np.nonzero(want[:m] != got[:m])[0]
    if len(d):
        print("  first diff at entry", d[0], want[d[0] - 2:d[0] + 3].tolist(), got[d[0] - 2:d[0] + 3].tolist())
d_in = torch.from_numpy(packed).cuda()
p = ffi.MzParams(); L.mz_params_nthash(C.byref(p), k, w, 0, 1)
ctx = sm.Context()
for c in range((nwin + chunk - 1) // chunk):
    wb, we = c * chunk, min(nwin, (c + 1) * chunk)
    # oracle entries of this window range: a window range [wb, we) emits pos where the selection of window j != window j-1
    sub, _ = o.run(packed, wb - 1 if wb else 0, (we + k + w - 2) - (wb - 1 if wb else 0), pr, "stream")
    sub = sub.astype(np.int64) + (wb - 1 if wb else 0)
    cap = len(sub) + 1000
    dp = torch.empty(cap, dtype=torch.int32, device="cuda")
    out = ffi.MzOut(dp.data_ptr(), None, None, cap, 0)
    assert L.mz_run_device(ctx.handle, 0, C.byref(p), d_in.data_ptr(), 0, n, wb, we, C.byref(out)) == 0
    g2 = dp[:out.count].cpu().numpy().view(np.uint32).astype(np.int64)
    # the oracle sub-run always emits its first window; the shard only if it differs from window wb-1
    a, bb = set(sub.tolist()), set(g2.tolist())
    print("  device sub-range", c, wb, we, "count", out.count, "oracle-sub", len(sub), "missing", sorted(a - bb)[:5], "extra", sorted(bb - a)[:5])
