"""The N-device paths of the library (mz_run / mz_run_skip_ambiguous / mz_run_batch on a context
that spans several devices): chunks of windows (or reads) are dealt round-robin over the devices
of the context, every chunk a contiguous shard with a (k+w-2)-base halo and one seam window, and
the outputs land in ONE ordered, globally indexed array (north_star, SURVEY 8e;
seam rule src/collect.rs:252-272).

The device list of a context may name a device more than once, which gives every listed entry
its own streams and buffers: `[0, 0]` exercises exactly the N-device code on a single-GPU box
(so this file never skips), and every test also runs on all visible GPUs when there are >= 2."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _device_lists():
    import torch

    n = torch.cuda.device_count()
    lists = [[0, 0], [0, 0, 0]]
    if n >= 2:
        lists.append(list(range(n)))
    return lists


def _builder(sm, k, w, canonical, mode):
    f = {(False, 0): sm.minimizers, (True, 0): sm.canonical_minimizers,
         (False, 1): sm.closed_syncmers, (True, 1): sm.canonical_closed_syncmers,
         (False, 2): sm.open_syncmers, (True, 2): sm.canonical_open_syncmers}[(canonical, mode)]
    return f(k, w)


def test_tiny_inputs_on_many_devices(sm, oracle, monkeypatch):
    """nwin from 0 to ndev^2 + 3 with 1..3 windows per chunk: fewer windows than devices, empty
    trailing chunks, chunk seams at every window (shard bounds were not clamped)."""
    monkeypatch.setenv("MZ_PIPELINE_MIN_WINDOWS", "1")
    packed = oracle.synth_packed(21, 4096)
    for devs in _device_lists():
        ctx = sm.Context(devs)
        nd = len(devs)
        for chunk in ("1", "2", "3"):
            monkeypatch.setenv("MZ_CHUNK_WINDOWS", chunk)
            for (k, w, canonical, mode) in ((31, 19, True, 0), (5, 3, False, 0), (7, 5, True, 1), (4, 3, False, 2)):
                l = k + w - 1
                b = _builder(sm, k, w, canonical, mode).context(ctx)
                pr = oracle.make_params(k, w, canonical=canonical, mode=mode)
                for nwin in list(range(0, nd * nd + 4)) + [97]:
                    n = l - 1 + nwin
                    want, wsk = oracle.run(packed, 1, n, pr, "stream", want_sk=(mode == 0))
                    pos, sk = sm.U32Vec(), sm.U32Vec()
                    (b.super_kmers(sk) if mode == 0 else b).run(sm.PackedSeq(packed, 1, n), pos)
                    assert np.array_equal(pos.array, want), (devs, chunk, k, w, mode, nwin)
                    if mode == 0:
                        assert np.array_equal(sk.array, wsk), (devs, chunk, k, w, nwin)
        ctx.close()


def test_multi_device_pipeline_matches_oracle(sm, oracle, monkeypatch):
    """30 Mbp through the default multi-device chunking (>= 6 chunks per device), fused values and
    super-k-mer starts, transfer codec on and off, skip-ambiguous with N runs across chunk seams."""
    n = 30_000_001
    off = 3
    packed = oracle.synth_packed(13, n + 8)
    rng = np.random.default_rng(4)
    bits = np.zeros(n + 8, dtype=np.uint8)
    for _ in range(30):
        a = int(rng.integers(0, n))
        bits[a:a + int(rng.integers(1, 200_000))] = 1
    for i in range(1, 16):  # ambiguous stretches across likely chunk boundaries
        bits[i * (n // 16) - 30:i * (n // 16) + 10] = 1
    amb = np.zeros((n + 8 + 7) // 8 + 16, dtype=np.uint8)
    pk = np.packbits(bits, bitorder="little")
    amb[:pk.size] = pk
    seq = sm.PackedSeq(packed, off, n)
    nseq = sm.PackedNSeq(seq, sm.BitSeq(amb, off, n))
    cases = []
    for (k, w, mode) in ((31, 19, 0), (21, 11, 1)):
        pr = oracle.make_params(k, w, canonical=True, mode=mode)
        epos, esk = oracle.run(packed, off, n, pr, "stream", want_sk=(mode == 0))
        evals = oracle.values_u64(packed, off, k, True, epos) if mode == 0 else None
        eamb = oracle.run_skip_ambiguous(packed, off, n, amb, off, pr)
        cases.append((k, w, mode, epos, esk, evals, eamb))
    for devs in _device_lists()[1:]:
        ctx = sm.Context(devs)
        assert ctx.device_count() == len(devs)
        for codec in ("on", "off"):
            if codec == "on":
                monkeypatch.setenv("MZ_DELTA_MAX_DEVICES", "64")
                monkeypatch.delenv("MZ_NO_POS_DELTA", raising=False)
            else:
                monkeypatch.setenv("MZ_NO_POS_DELTA", "1")
            for (k, w, mode, epos, esk, evals, eamb) in cases:
                b = _builder(sm, k, w, True, mode).context(ctx)
                if mode == 0:
                    pos, sk, vals = b.super_kmers(sm.U32Vec()).run_with_values(seq, 64)
                    assert np.array_equal(pos, epos) and np.array_equal(sk, esk), (devs, codec)
                    assert np.array_equal(vals, evals), (devs, codec)
                    v = sm.U32Vec()
                    out = b.run(seq, v)
                    assert np.array_equal(v.array, epos)
                    assert np.array_equal(out.values_u64(), evals)
                else:
                    assert np.array_equal(b.run_once(seq), epos), (devs, codec)
                assert np.array_equal(b.run_skip_ambiguous_windows_once(nseq), eamb), (devs, codec, mode)
        t = ctx.last_timing()
        assert t["kernel_launches"] >= 2  # more than one chunk
        ctx.close()


def test_capacity_error_and_reuse_on_many_devices(sm, oracle, monkeypatch):
    """A too small output buffer reports the needed size (MZ_ERR_CAPACITY) from the pipelined
    N-device path and leaves the context usable."""
    import ctypes as C
    from importlib import import_module

    ffi = import_module("simd-minimizers_b200._ffi")
    L = ffi.lib()
    monkeypatch.setenv("MZ_PIPELINE_MIN_WINDOWS", "1")
    monkeypatch.setenv("MZ_CHUNK_WINDOWS", "50000")
    n = 1_000_000
    packed = oracle.synth_packed(2, n + 4)
    pr = oracle.make_params(31, 19, canonical=True)
    want, _ = oracle.run(packed, 0, n, pr, "stream")
    ctx = sm.Context([0, 0])
    p = ffi.MzParams()
    L.mz_params_nthash(C.byref(p), 31, 19, 0, 1)
    pos = np.zeros(len(want), dtype=np.uint32)
    out = ffi.MzOut(pos.ctypes.data, None, None, 1000, 0)
    rc = L.mz_run(ctx.handle, C.byref(p), packed.ctypes.data, 0, n, C.byref(out))
    assert rc == ffi.MZ_ERR_CAPACITY and out.count == len(want)
    out = ffi.MzOut(pos.ctypes.data, None, None, len(want), 0)
    assert L.mz_run(ctx.handle, C.byref(p), packed.ctypes.data, 0, n, C.byref(out)) == 0
    assert out.count == len(want) and np.array_equal(pos, want)
    ctx.close()


def test_batch_reads_sharded_over_devices(sm, oracle, monkeypatch):
    """BASELINE config 5's shape at test size: 1.2 M x 150 bp reads at a 38-byte stride, canonical
    k=21 w=11, dealt over the devices of the context chunk by chunk; every read must give exactly
    what the per-read reference loop gives (bench/src/bin/paper.rs:98-105), CSR offsets global."""
    n_reads, read_len, stride = 1_200_000, 150, 38
    k, w = 21, 11
    packed = oracle.synth_packed(99, n_reads * stride * 4 + 64)
    pr = oracle.make_params(k, w, canonical=True)
    eo, ep, esk, ev = oracle.run_reads(packed, n_reads, stride, read_len, pr, threads=8, want_sk=True, want_val=True)
    for devs in _device_lists():
        ctx = sm.Context(devs)
        for chunk in (None, "100003"):
            if chunk:
                monkeypatch.setenv("MZ_BATCH_CHUNK_READS", chunk)
            else:
                monkeypatch.delenv("MZ_BATCH_CHUNK_READS", raising=False)
            b = sm.canonical_minimizers(k, w).context(ctx).super_kmers(sm.U32Vec())
            offs, pos, sk, vals = b.run_batch(packed, stride_bytes=stride, read_len=read_len, n_reads=n_reads)
            assert np.array_equal(offs, eo), (devs, chunk)
            assert np.array_equal(pos, ep) and np.array_equal(sk, esk), (devs, chunk)
            assert np.array_equal(vals, ev), (devs, chunk)
        # ragged reads in storage order stream through the same chunks
        monkeypatch.setenv("MZ_BATCH_CHUNK_READS", "70001")
        rng = np.random.default_rng(8)
        m = 200_000
        lens = rng.integers(0, 400, size=m).astype(np.uint32)
        starts = np.concatenate([[0], np.cumsum(lens.astype(np.uint64) + rng.integers(0, 9, size=m).astype(np.uint64))[:-1]]).astype(np.uint64)
        offs, pos, _, _ = sm.canonical_minimizers(k, w).context(ctx).run_batch(packed, starts=starts, lens=lens, value_bits=0)
        for r in (0, 1, 17, m // 2, m - 1):
            want, _ = oracle.run(packed, int(starts[r]), int(lens[r]), pr)
            assert np.array_equal(pos[int(offs[r]):int(offs[r + 1])], want), (devs, r)
        tot = sum(max(0, int(x) - (k + w - 1) + 1) for x in lens[:2000])
        assert int(offs[2000]) <= tot
        ctx.close()


def test_batch_into_pinned_arrays(sm, oracle, monkeypatch):
    """Pinned caller arrays (mz_host_alloc): positions AND the CSR offsets are copied straight from
    the devices to their final place (offsets rebased on the device, no bounce buffer, no host
    pass) -- the path bench.py's config 5 takes.  Checked against the per-read oracle loop."""
    import ctypes as C
    import importlib

    ffi = importlib.import_module("simd-minimizers_b200._ffi")
    L = ffi.lib()
    n_reads, read_len, stride = 300_000, 150, 38
    k, w = 21, 11
    packed = oracle.synth_packed(5, n_reads * stride * 4 + 64)
    pr = oracle.make_params(k, w, canonical=True)
    eo, ep, _, _ = oracle.run_reads(packed, n_reads, stride, read_len, pr, threads=8)
    nbytes = n_reads * stride
    cap = len(ep) + 1024

    def pinned(n, dtype):
        q = C.c_void_p()
        assert L.mz_host_alloc(C.byref(q), n * np.dtype(dtype).itemsize) == 0
        arr = np.ctypeslib.as_array(C.cast(q, C.POINTER(C.c_uint8)), shape=(n * np.dtype(dtype).itemsize,)).view(dtype)
        return q, arr

    q_in, h_in = pinned(nbytes + 64, np.uint8)
    h_in[:nbytes] = packed.view(np.uint8)[:nbytes]
    q_off, h_off = pinned(n_reads + 1, np.uint64)
    q_pos, h_pos = pinned(cap, np.uint32)
    p = ffi.MzParams()
    L.mz_params_nthash(C.byref(p), k, w, 0, 1)
    monkeypatch.setenv("MZ_BATCH_CHUNK_READS", "40009")   # 8 chunks: bases of all sizes to add
    try:
        for devs in _device_lists():
            ctx = sm.Context(devs)
            h_off[:] = 0xDEADBEEF
            h_pos[:] = 0xFFFFFFFF
            out = ffi.MzOut(q_pos.value, None, None, cap, 0)
            rc = L.mz_run_batch(ctx.handle, C.byref(p), q_in.value, nbytes, n_reads, None, None, stride, read_len,
                                q_off.value, C.byref(out))
            assert rc == 0, rc
            assert out.count == len(ep)
            assert np.array_equal(h_off, eo), devs
            assert np.array_equal(h_pos[:len(ep)], ep), devs
            ctx.close()
    finally:
        for q in (q_in, q_off, q_pos):
            L.mz_host_free(q)


def test_pcie_probe(sm):
    """mz_pcie_probe: the measured copy ceiling bench.py reports next to e2e."""
    import ctypes as C
    from importlib import import_module

    ffi = import_module("simd-minimizers_b200._ffi")
    for devs in _device_lists()[::2]:
        ctx = sm.Context(devs)
        r = ffi.MzPcieResult()
        ffi.check(ffi.lib().mz_pcie_probe(ctx.handle, 64 << 20, 3, C.byref(r)))
        assert r.n_devices == len(devs)
        assert r.h2d_gbs > 1 and r.d2h_gbs > 1 and r.bidir_h2d_gbs > 0.5 and r.bidir_d2h_gbs > 0.5
        ctx.close()
