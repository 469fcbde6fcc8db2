"""GPU parity for run_skip_ambiguous_windows (src/lib.rs:451-496) through the C ABI: the CUDA
path against the oracle's restatement of the reference stream + SKIP_MAX collector
(src/minimizers.rs:169-214, src/intrinsics/dedup.rs:147-155, src/syncmers.rs:152), and the
reference's own property test (src/test.rs:429-482): no k-mer of the output holds an ambiguous
base."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ascii_with_n(rng, n, frac, runs, maxrun=300):
    s = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
    k = int(n * frac)
    if k:
        s[rng.integers(0, n, size=k)] = ord("N")
    for _ in range(runs):
        a = int(rng.integers(0, n))
        s[a:a + int(rng.integers(1, maxrun))] = ord("N")
    return s.tobytes()


def _builder(sm, k, w, mode):
    return {0: sm.canonical_minimizers, 1: sm.canonical_closed_syncmers,
            2: sm.canonical_open_syncmers}[mode](k, w)


def _check(sm, oracle, packed, amb, off, n, k, w, mode, kind="nt"):
    nseq = sm.PackedNSeq(sm.PackedSeq(packed, off, n), sm.BitSeq(amb, off, n))
    b = _builder(sm, k, w, mode)
    if kind == "mul":
        b = b.hasher(sm.MulHasher(k, True))
    pr = oracle.make_params(k, w, canonical=True, mode=mode, hasher=oracle.make_hasher(kind, True))
    want = oracle.run_skip_ambiguous(packed, off, n, amb, off, pr, "stream")
    pos = sm.U32Vec()
    out = b.run_skip_ambiguous_windows(nseq, pos)
    tag = (k, w, n, off, mode, kind)
    assert np.array_equal(pos.array, want), tag
    length = k if mode == 0 else k + w - 1
    if length <= 32:
        wv = oracle.values_u64(packed, off, length, True, want)
        assert np.array_equal(out.values_u64(), wv), tag
        fpos, _, fval = b.run_with_values(nseq, 64, skip_ambiguous=True)
        assert np.array_equal(fpos, want) and np.array_equal(fval, wv), tag
    elif length <= 64:
        wv = oracle.values_u128(packed, off, length, True, want)
        assert np.array_equal(out._values(128), wv), tag
        fpos, _, fval = b.run_with_values(nseq, 128, skip_ambiguous=True)
        assert np.array_equal(fpos, want) and np.array_equal(fval, wv), tag
    return want


def test_skip_ambiguous_reference_property_and_grid(sm, oracle):
    """The reference's test: 100 bases, ~1 % N, every (k, w) with odd l <= 64."""
    rng = np.random.default_rng(11)
    ascii_ = _ascii_with_n(rng, 100, 0.01, 1, maxrun=4)
    packed, amb = oracle.pack_ascii_n(ascii_)
    isn = np.frombuffer(ascii_, dtype=np.uint8) == ord("N")
    assert isn.any()
    for k in range(1, 65):
        for w in range(1, 64):
            if (k + w - 1) % 2 == 0 or k + w - 1 > 64:
                continue
            got = _check(sm, oracle, packed, amb, 0, 100, k, w, 0)
            for p in got.tolist():
                assert p != oracle.SKIPPED and not isn[p:p + k].any()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_skip_ambiguous_grid(sm, oracle, mode):
    rng = np.random.default_rng(100 + mode)
    n_max = 3000
    ascii_ = _ascii_with_n(rng, n_max + 16, 0.004, 6, maxrun=150)
    packed, amb = oracle.pack_ascii_n(ascii_)
    ks = [1, 2, 5, 21, 31, 32, 33, 64, 65]
    ws = [1, 2, 3, 5, 11, 19, 31, 32, 33, 64, 101]
    for k in ks:
        for w in ws:
            if (k + w - 1) % 2 == 0:
                continue
            if mode == 2 and w % 2 == 0:
                continue
            for (off, n) in ((0, n_max), (3, 997), (9, 2 * (k + w)), (1, k + w - 1), (2, k + w - 2)):
                _check(sm, oracle, packed, amb, off, n, k, w, mode, "mul" if (k + w) % 3 == 0 else "nt")


def test_skip_ambiguous_edge_masks(sm, oracle):
    """All clean == plain run; all ambiguous == empty; single N; N at both ends; long N runs."""
    rng = np.random.default_rng(7)
    n = 50_000
    base = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
    for k, w in ((31, 19), (21, 11), (5, 3), (15, 41)):
        # all clean
        packed, amb = oracle.pack_ascii_n(base.tobytes())
        assert not amb.any()
        got = _check(sm, oracle, packed, amb, 0, n, k, w, 0)
        plain = sm.canonical_minimizers(k, w).run_once(sm.PackedSeq(packed, 0, n))
        assert np.array_equal(got, plain)
        # all ambiguous
        packed, amb = oracle.pack_ascii_n(b"N" * n)
        assert len(_check(sm, oracle, packed, amb, 0, n, k, w, 0)) == 0
        # single N in the middle, N at both ends, N runs around tile / thread seams
        for marks in ([n // 2], [0, n - 1], list(range(9216 - 3, 9216 + 40)), list(range(288 * 7, 288 * 7 + 500)),
                      list(range(0, n, 97))):
            s = base.copy()
            s[marks] = ord("n")
            packed, amb = oracle.pack_ascii_n(s.tobytes())
            for mode in (0, 1, 2):
                if mode == 2 and w % 2 == 0:
                    continue
                _check(sm, oracle, packed, amb, 0, n, k, w, mode)


def test_pack_ascii_n_and_ascii_run(sm, oracle):
    rng = np.random.default_rng(3)
    for n in (0, 1, 31, 32, 33, 100, 4097, 1_000_003):
        s = rng.choice(np.frombuffer(b"ACGTacgtNnRYKM-*", dtype=np.uint8), size=n).tobytes()
        want_p, want_a = oracle.pack_ascii_n(s)
        v = sm.PackedNSeqVec.from_ascii(s)
        assert np.array_equal(v.seq.data[:(n + 3) // 4], want_p[:(n + 3) // 4])
        assert np.array_equal(v.amb[:(n + 7) // 8], want_a[:(n + 7) // 8])
        if n >= 100:
            pr = oracle.make_params(21, 11, canonical=True)
            want = oracle.run_skip_ambiguous(want_p, 0, n, want_a, 0, pr)
            got = sm.canonical_minimizers(21, 11).run_skip_ambiguous_windows_once(sm.AsciiSeq(s))
            assert np.array_equal(got, want)
            got = sm.canonical_minimizers(21, 11).run_skip_ambiguous_windows_once(v)
            assert np.array_equal(got, want)


def test_skip_ambiguous_errors_and_append(sm, oracle):
    rng = np.random.default_rng(5)
    s = _ascii_with_n(rng, 2000, 0.01, 2)
    v = sm.PackedNSeqVec.from_ascii(s)
    with pytest.raises(TypeError):  # forward builders have no run_skip_ambiguous_windows
        sm.minimizers(5, 3).run_skip_ambiguous_windows(v, sm.U32Vec())
    with pytest.raises(TypeError):  # nor do builders with super-k-mers
        sm.canonical_minimizers(5, 3).super_kmers(sm.U32Vec()).run_skip_ambiguous_windows(v, sm.U32Vec())
    with pytest.raises(AssertionError):  # even l
        sm.canonical_minimizers(4, 3).run_skip_ambiguous_windows(v, sm.U32Vec())
    # C ABI: forward params are refused
    from importlib import import_module
    ffi = import_module("simd-minimizers_b200._ffi")
    L = ffi.lib()
    p = ffi.MzParams()
    L.mz_params_nthash(C.byref(p), 5, 3, 0, 0)
    pos = np.zeros(4096, dtype=np.uint32)
    out = ffi.MzOut(pos.ctypes.data, None, None, 4096, 0)
    rc = L.mz_run_skip_ambiguous(sm.default_context().handle, C.byref(p), v.seq.data.ctypes.data, 0, 2000,
                                 v.amb.ctypes.data, 0, C.byref(out))
    assert rc == 6  # MZ_ERR_NOT_CANONICAL
    # append semantics as for run(): positions are appended
    pos = sm.U32Vec()
    b = sm.canonical_minimizers(7, 5)
    b.run_skip_ambiguous_windows(v, pos)
    first = pos.array.copy()
    b.run_skip_ambiguous_windows(v, pos)
    assert np.array_equal(pos.array, np.concatenate([first, first]))


def test_skip_ambiguous_large_pipelined_and_device(sm, oracle, monkeypatch):
    """40 Mbp with genome-like N runs: chunk-pipelined host path (seams inside N runs and inside
    clean stretches), the device-resident entry point on window sub-ranges, values."""
    import torch
    from importlib import import_module
    ffi = import_module("simd-minimizers_b200._ffi")
    n = 40_000_000
    packed = oracle.synth_packed(9, n + 5)
    rng = np.random.default_rng(2)
    bits = np.zeros(n + 5, dtype=np.uint8)
    for _ in range(40):
        a = int(rng.integers(0, n))
        bits[a:a + int(rng.integers(1, 400_000))] = 1
    bits[rng.integers(0, n, size=2000)] = 1
    amb = np.zeros((n + 5 + 7) // 8 + 16, dtype=np.uint8)
    pk = np.packbits(bits, bitorder="little")
    amb[:pk.size] = pk
    k, w, off = 31, 19, 5
    nn = n - 3
    pr = oracle.make_params(k, w, canonical=True)
    want = oracle.run_skip_ambiguous(packed, off, nn, amb, off, pr)
    wantv = oracle.values_u64(packed, off, k, True, want)
    nseq = sm.PackedNSeq(sm.PackedSeq(packed, off, nn), sm.BitSeq(amb, off, nn))
    for chunk in (None, "3000017"):
        if chunk:
            monkeypatch.setenv("MZ_CHUNK_WINDOWS", chunk)
            monkeypatch.setenv("MZ_PIPELINE_MIN_WINDOWS", "1000")
        pos = sm.U32Vec()
        out = sm.canonical_minimizers(k, w).run_skip_ambiguous_windows(nseq, pos)
        assert np.array_equal(pos.array, want)
        assert np.array_equal(out.values_u64(), wantv)
        fpos, _, fval = sm.canonical_minimizers(k, w).run_with_values(nseq, 64, skip_ambiguous=True)
        assert np.array_equal(fpos, want) and np.array_equal(fval, wantv)
    monkeypatch.delenv("MZ_CHUNK_WINDOWS")
    monkeypatch.delenv("MZ_PIPELINE_MIN_WINDOWS")
    # closed syncmers on the same input
    prs = oracle.make_params(k, 11, canonical=True, mode=1)
    wants = oracle.run_skip_ambiguous(packed, off, nn, amb, off, prs)
    got = sm.canonical_closed_syncmers(k, 11).run_skip_ambiguous_windows_once(nseq)
    assert np.array_equal(got, wants)
    # device-resident, window sub-ranges concatenate to the whole
    L = ffi.lib()
    ctx = sm.default_context()
    d_seq = torch.from_numpy(packed).cuda()
    d_amb = torch.from_numpy(amb).cuda()
    p = ffi.MzParams()
    L.mz_params_nthash(C.byref(p), k, w, 0, 1)
    p.value_bits = 64
    nwin = nn - (k + w - 1) + 1
    cuts = [0, 1, 12345, nwin // 3, nwin // 3 + 1, 2 * nwin // 3, nwin]
    parts, vparts = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        cap = b - a
        d_pos = torch.empty(cap, dtype=torch.int32, device="cuda")
        d_val = torch.empty(cap, dtype=torch.int64, device="cuda")
        out = ffi.MzOut(d_pos.data_ptr(), None, d_val.data_ptr(), cap, 0)
        ffi.check(L.mz_run_device_skip_ambiguous(ctx.handle, 0, C.byref(p), d_seq.data_ptr(), off, nn,
                                                 d_amb.data_ptr(), off, a, b, C.byref(out)))
        parts.append(d_pos[:out.count].cpu().numpy().view(np.uint32))
        vparts.append(d_val[:out.count].cpu().numpy().view(np.uint64))
    got = np.concatenate(parts)
    # a sub-range always emits its first clean window when the window before it is ambiguous or
    # outside the range start 0; seams inside a clean run dedup against the seam window
    assert np.array_equal(got, want)
    assert np.array_equal(np.concatenate(vparts), wantv)


def test_in_process_multi_device_shards(sm, oracle):
    """mz_run / mz_run_skip_ambiguous with a context that spans every visible GPU: windows are cut
    into contiguous shards (halo k+w-2 bases + one seam window), outputs gathered in order.
    Needs >= 2 GPUs (`gpurun --gpus 2`); skipped on a single-GPU box."""
    import torch
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs at least two GPUs")
    ctx = sm.Context(list(range(ndev)))
    assert ctx.device_count() == ndev
    n = 30_000_001
    packed = oracle.synth_packed(13, n + 8)
    rng = np.random.default_rng(4)
    bits = np.zeros(n + 8, dtype=np.uint8)
    for _ in range(30):
        a = int(rng.integers(0, n))
        bits[a:a + int(rng.integers(1, 200_000))] = 1
    # ambiguous stretches across every shard boundary
    per = (n - 49 + 1 + ndev - 1) // ndev
    for i in range(1, ndev):
        bits[i * per - 30:i * per + 10] = 1
    amb = np.zeros((n + 8 + 7) // 8 + 16, dtype=np.uint8)
    pk = np.packbits(bits, bitorder="little")
    amb[:pk.size] = pk
    off = 3
    for (k, w, mode) in ((31, 19, 0), (21, 11, 1)):
        pr = oracle.make_params(k, w, canonical=True, mode=mode)
        b = {0: sm.canonical_minimizers, 1: sm.canonical_closed_syncmers}[mode](k, w).context(ctx)
        # plain run (+ super-k-mer starts for minimizers)
        epos, esk = oracle.run(packed, off, n, pr, "stream", want_sk=(mode == 0))
        pos, sk = sm.U32Vec(), sm.U32Vec()
        out = (b.super_kmers(sk) if mode == 0 else b).run(sm.PackedSeq(packed, off, n), pos)
        assert np.array_equal(pos.array, epos)
        if mode == 0:
            assert np.array_equal(sk.array, esk)
            assert np.array_equal(out.values_u64(), oracle.values_u64(packed, off, k, True, epos))
        # skip-ambiguous
        want = oracle.run_skip_ambiguous(packed, off, n, amb, off, pr)
        nseq = sm.PackedNSeq(sm.PackedSeq(packed, off, n), sm.BitSeq(amb, off, n))
        got = b.run_skip_ambiguous_windows_once(nseq)
        assert np.array_equal(got, want)
    ctx.close()
