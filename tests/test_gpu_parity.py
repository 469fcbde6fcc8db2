"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI,
against the CPU oracle -- bit-exact positions, super-k-mer starts, values and syncmers on the
reference's (k, w, len, offset) grid (src/test.rs:24-51), the reference's golden vectors, and
size-independent properties at larger sizes."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")
KS = [1, 2, 3, 4, 5, 31, 32, 33, 63, 64, 65]
WS = [1, 2, 3, 4, 5, 31, 32, 33, 63, 64, 65]


def _builder(sm, k, w, canonical, mode):
    f = {(False, 0): sm.minimizers, (True, 0): sm.canonical_minimizers,
         (False, 1): sm.closed_syncmers, (True, 1): sm.canonical_closed_syncmers,
         (False, 2): sm.open_syncmers, (True, 2): sm.canonical_open_syncmers}[(canonical, mode)]
    return f(k, w)


def _check_case(sm, oracle, packed, off, n, k, w, canonical, mode, kind="nt", hash_canon=None):
    if hash_canon is None:
        hash_canon = canonical
    seq = sm.PackedSeq(packed, off, n)
    b = _builder(sm, k, w, canonical, mode)
    if kind != "nt" or hash_canon != canonical:
        H = sm.NtHasher if kind == "nt" else sm.MulHasher
        b = b.hasher(H(k, hash_canon))
    pr = oracle.make_params(k, w, canonical=canonical, mode=mode,
                            hasher=oracle.make_hasher(kind, hash_canon))
    epos, esk = oracle.run(packed, off, n, pr, "stream", want_sk=(mode == 0))
    pos, sk = sm.U32Vec(), sm.U32Vec()
    out = (b.super_kmers(sk) if mode == 0 else b).run(seq, pos)
    tag = (k, w, n, off, canonical, mode, kind)
    assert np.array_equal(pos.array, epos), tag
    if mode == 0:
        assert np.array_equal(sk.array, esk), tag
    length = k if mode == 0 else k + w - 1
    # values twice: the lazy Output iterator (mz_values) and fused into the run (value_bits)
    if length <= 32:
        want = oracle.values_u64(packed, off, length, canonical, epos)
        assert np.array_equal(out.values_u64(), want), tag
        fpos, fsk, fval = (b.super_kmers(sm.U32Vec()) if mode == 0 else b).run_with_values(seq, 64)
        assert np.array_equal(fpos, epos) and np.array_equal(fval, want), tag
        assert mode != 0 or np.array_equal(fsk, esk), tag
    elif length <= 64:
        want = oracle.values_u128(packed, off, length, canonical, epos)
        assert np.array_equal(out._values(128), want), tag
        fpos, _, fval = b.run_with_values(seq, 128)
        assert np.array_equal(fpos, epos) and np.array_equal(fval, want), tag


def test_reference_golden_vectors(sm):
    v = json.load(open(GOLDEN))
    d = v["doc_forward"]
    s = sm.PackedSeqVec.from_ascii(d["seq"].encode())
    assert sm.minimizer_positions(s, d["k"], d["w"]).tolist() == d["pos"]
    d = v["doc_canonical"]
    s = sm.PackedSeqVec.from_ascii(d["seq"].encode())
    assert sm.canonical_minimizer_positions(s, d["k"], d["w"]).tolist() == d["pos"]
    pos = sm.U32Vec()
    vals = sm.canonical_minimizers(d["k"], d["w"]).run(s, pos).values_u64()
    assert pos.tolist() == d["pos"] and vals.tolist() == d["values"]
    rc = s.as_slice().to_revcomp()
    rpos = sm.U32Vec()
    rvals = sm.canonical_minimizers(d["k"], d["w"]).run(rc, rpos).values_u64()
    assert rpos.tolist() == d["rc_pos"] and rvals.tolist()[::-1] == d["values"]


def test_closed_syncmer_values_all_g(sm):
    n = 100                                     # src/test.rs:578-597
    s = sm.PackedSeqVec.from_ascii(b"G" * n)
    for k in range(1, 10):
        for w in range(1, 10):
            pos = sm.U32Vec()
            vals = sm.closed_syncmers(k, w).run(s, pos).values_u64()
            assert len(vals) == n - (k + w - 1) + 1
            assert (vals == (1 << (2 * (k + w - 1))) - 1).all()


@pytest.mark.parametrize("canonical", [False, True])
def test_minimizer_grid(sm, oracle, canonical):
    """src/test.rs:55-110, 155-277 grid: every (k, w) pair, short + long lengths, offsets 0..3."""
    rng = np.random.default_rng(11 + canonical)
    base = oracle.synth_packed(99, 8192 + 8)
    ks = KS + [int(x) for x in rng.integers(6, 100, 3)]
    ws = WS + [int(x) for x in rng.integers(6, 100, 3)]
    for k in ks:
        for w in ws:
            if canonical and (k + w - 1) % 2 == 0:
                continue
            lens = [int(x) for x in rng.integers(0, 100, 3)] + [k + w - 2, k + w - 1, k + w,
                    int(rng.integers(100, 8192))]
            for n in lens:
                off = int(rng.integers(0, 4))
                _check_case(sm, oracle, base, off, n, k, w, canonical, 0)


def test_all_short_lengths(sm, oracle):
    base = oracle.synth_packed(5, 128)
    for n in range(0, 100):
        for (k, w, c) in ((5, 7, True), (3, 4, False), (31, 19, True), (1, 1, True)):
            _check_case(sm, oracle, base, n % 4, n, k, w, c, 0)


@pytest.mark.parametrize("mode", [1, 2])
def test_syncmer_grid(sm, oracle, mode):
    """src/test.rs:519-574, 600-639"""
    rng = np.random.default_rng(100 + mode)
    base = oracle.synth_packed(123, 8192 + 8)
    for k in KS[:8] + [int(x) for x in rng.integers(6, 60, 2)]:
        for w in WS[:8] + [int(x) for x in rng.integers(6, 60, 2)]:
            if mode == 2 and w % 2 == 0:
                continue
            for canonical in (False, True):
                if canonical and (k + w - 1) % 2 == 0:
                    continue
                for n in (int(rng.integers(0, 100)), int(rng.integers(100, 4096))):
                    _check_case(sm, oracle, base, int(rng.integers(0, 4)), n, k, w, canonical, mode)


def test_hashers(sm, oracle):
    """MulHasher, and a forward builder with a canonical hasher (src/minimizers.rs:69-71)."""
    rng = np.random.default_rng(5)
    base = oracle.synth_packed(77, 6000)
    for _ in range(40):
        k, w = int(rng.integers(1, 50)), int(rng.integers(1, 40))
        n, off = int(rng.integers(0, 5000)), int(rng.integers(0, 4))
        _check_case(sm, oracle, base, off, n, k, w, False, 0, kind="mul")
        _check_case(sm, oracle, base, off, n, k, w, False, 0, kind="nt", hash_canon=True)
        if (k + w - 1) % 2 == 1:
            _check_case(sm, oracle, base, off, n, k, w, True, 0, kind="mul")


def test_degenerate_sequences(sm, oracle):
    """Homopolymers / short periods: every window emits (density 1) or ties everywhere."""
    for pattern in (b"A", b"G", b"AC", b"ACGT", b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAT"):
        s = (pattern * (3000 // len(pattern) + 1))[:3000]
        packed = oracle.pack_ascii(s)
        for (k, w, c) in ((31, 19, True), (21, 11, False), (5, 3, True), (8, 1, False)):
            for mode in (0, 1):
                _check_case(sm, oracle, packed, 0, len(s), k, w, c, mode)


def test_error_behaviour(sm):
    s = sm.PackedSeqVec.random(200, 1)
    with pytest.raises(AssertionError, match="must be odd"):
        sm.canonical_minimizer_positions(s, 4, 3)
    with pytest.raises(AssertionError, match="Open syncmers require odd"):
        sm.open_syncmers(5, 4).run_once(s)
    with pytest.raises(AssertionError):
        sm.canonical_minimizers(5, 7).hasher(sm.NtHasher(5, False)).run_once(s)
    pos = sm.U32Vec()
    out = sm.canonical_minimizers(33, 3).run(s, pos)
    with pytest.raises(AssertionError):
        out.values_u64()
    assert len(out.values_u128()) == len(pos)


def test_append_semantics(sm, oracle):
    """Positions are appended; the first new element is dropped if it repeats the last one
    (src/collect.rs:257,267)."""
    s = sm.PackedSeqVec.random(500, 3)
    first = sm.minimizer_positions(s, 7, 5)
    v = sm.U32Vec()
    sm.minimizers(7, 5).run(s, v)
    sm.minimizers(7, 5).run(s, v)
    assert v.tolist() == first.tolist() + first.tolist()
    v = sm.U32Vec([int(first[0])])
    sm.minimizers(7, 5).run(s, v)
    assert v.tolist() == first.tolist()


def test_lazy_values_cover_the_whole_vector(sm, oracle):
    """Output::values_* iterate ALL of min_pos (src/lib.rs:598-612), including positions appended
    by earlier runs; positions whose k-mer runs past the sequence are an error (read_kmer asserts)."""
    n, k, w = 5000, 21, 11
    packed = oracle.synth_packed(5, n + 4)
    seq = sm.PackedSeq(packed, 1, n)
    v = sm.U32Vec([7, 100])
    out = sm.canonical_minimizers(k, w).run(seq, v)
    assert v.tolist()[:2] == [7, 100] and len(v) > 2
    assert np.array_equal(out.values_u64(), oracle.values_u64(packed, 1, k, True, v.array))
    got = out.values_u128()
    want = oracle.values_u128(packed, 1, k, True, v.array)
    assert got == [int(lo) | (int(hi) << 64) for lo, hi in want]
    pv = out.pos_and_values_u64()
    assert [a for a, _ in pv] == v.tolist()
    bad = sm.U32Vec([n - k + 1])
    with pytest.raises(sm.MzError):
        sm.Output(sm.canonical_minimizers(k, w), seq, bad).values_u64()
    # forward builder: plain k-mers; syncmers: l-mers
    v2 = sm.U32Vec()
    o2 = sm.closed_syncmers(9, 5).run(seq, v2)
    assert np.array_equal(o2.values_u64(), oracle.values_u64(packed, 1, 13, False, v2.array))


def test_run_scalar_overwrites(sm, oracle):
    """run_scalar / run_scalar_once (src/lib.rs:358-378, 504-533): same results, but the scalar
    collectors overwrite the vectors from index 0 (src/collect.rs:15-76) instead of appending."""
    n = 3000
    packed = oracle.synth_packed(77, n + 4)
    seq = sm.PackedSeq(packed, 2, n)
    for (k, w, canonical, mode) in ((21, 11, True, 0), (7, 5, False, 0), (9, 5, False, 1), (9, 5, True, 2)):
        b = _builder(sm, k, w, canonical, mode)
        pr = oracle.make_params(k, w, canonical=canonical, mode=mode)
        epos, esk = oracle.run(packed, 2, n, pr, "stream", want_sk=(mode == 0))
        assert np.array_equal(b.run_scalar_once(seq), epos)
        v = sm.U32Vec([1, 2, 3])
        out = b.run_scalar(seq, v)
        assert np.array_equal(v.array, epos)
        assert len(out.values_u64()) == len(epos)
        if mode == 0:
            sk = sm.U32Vec([9, 9])
            v = sm.U32Vec([4])
            b.super_kmers(sk).run_scalar(seq, v)
            assert np.array_equal(v.array, epos) and np.array_equal(sk.array, esk)
            assert np.array_equal(b.super_kmers(sm.U32Vec()).run_scalar_once(seq), epos)
    # too short: output cleared, index vector untouched (src/collect.rs:45-48)
    v, sk = sm.U32Vec([5]), sm.U32Vec([6])
    sm.minimizers(7, 5).super_kmers(sk).run_scalar(sm.PackedSeq(packed, 0, 8), v)
    assert v.tolist() == [] and sk.tolist() == [6]


def test_table_hashers(sm, oracle):
    """Hashers cross the ABI as their per-base tables (mz_params_set_tables): the built-in
    NtHasher tables passed explicitly give the built-in result; arbitrary (seeded-like) tables
    agree with the oracle run on the same tables."""
    n = 20000
    packed = oracle.synth_packed(3, n + 4)
    seq = sm.PackedSeq(packed, 0, n)
    F = [0x95c60474, 0x62a02b4c, 0x82572324, 0x4be24456]
    Cc = [F[b ^ 2] for b in range(4)]
    a = sm.canonical_minimizers(31, 19).run_once(seq)
    b = sm.canonical_minimizers(31, 19).hasher(sm.NtHasher.from_tables(31, F, Cc, 7, True)).run_once(seq)
    assert np.array_equal(a, b)
    rng = np.random.default_rng(11)
    for rot in (7, 1, 13):
        f = [int(x) for x in rng.integers(0, 2**32, 4)]
        c = [f[b ^ 2] for b in range(4)]
        for canonical in (False, True):
            h = oracle.make_hasher_tables(f, c, rot, canonical)
            pr = oracle.make_params(21, 11, canonical=canonical, hasher=h)
            want, _ = oracle.run(packed, 0, n, pr, "stream")
            got = (sm.canonical_minimizers if canonical else sm.minimizers)(21, 11).hasher(
                sm.NtHasher.from_tables(21, f, c, rot, canonical)).run_once(seq)
            assert np.array_equal(got, want), (rot, canonical)
    with pytest.raises(NotImplementedError):
        sm.MulHasher.new_with_seed(21, 1234)


def test_medium_exact_and_rc_symmetry(sm, oracle):
    """10 Mbp (BASELINE config 1 size) exact compare, plus rc symmetry at the same size."""
    n = 10_000_000
    packed = oracle.synth_packed(42, n)
    for (k, w, c) in ((21, 11, False), (31, 19, True)):
        _check_case(sm, oracle, packed, 0, n, k, w, c, 0)
    seq = sm.PackedSeq(packed, 0, n)
    rc = sm.PackedSeq(oracle.revcomp(packed, 0, n), 0, n)
    fpos, rpos = sm.U32Vec(), sm.U32Vec()
    fv = sm.canonical_minimizers(31, 19).run(seq, fpos).values_u64()
    rv = sm.canonical_minimizers(31, 19).run(rc, rpos).values_u64()
    assert len(fpos) == len(rpos)
    assert (fpos.array.astype(np.int64) + rpos.array[::-1].astype(np.int64) == n - 31).all()
    assert np.array_equal(fv, rv[::-1])


def test_device_resident_and_window_ranges(sm, oracle):
    """mz_run_device: device pointers in/out, arbitrary window sub-ranges concatenate to the
    full result (the multi-GPU shard rule)."""
    import ctypes as C
    import importlib

    import torch

    ffi = importlib.import_module("simd-minimizers_b200._ffi")
    L = ffi.lib()
    n, k, w = 300_000, 31, 19
    packed = oracle.synth_packed(9, n + 1)
    pr = oracle.make_params(k, w, canonical=True)
    epos, esk = oracle.run(packed, 1, n, pr, want_sk=True)
    evals = oracle.values_u64(packed, 1, k, True, epos)
    d_in = torch.from_numpy(packed).cuda()
    ctx = sm.Context()
    p = ffi.MzParams()
    L.mz_params_nthash(C.byref(p), k, w, 0, 1)
    p.want_sk, p.value_bits = 1, 64
    nwin = n - (k + w - 1) + 1
    cuts = [0, 1, 77, 4096, 150_001, nwin]
    got_p, got_s, got_v = [], [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        cap = b - a
        dp = torch.empty(cap, dtype=torch.int32, device="cuda")
        ds = torch.empty(cap, dtype=torch.int32, device="cuda")
        dv = torch.empty(cap, dtype=torch.int64, device="cuda")
        out = ffi.MzOut(dp.data_ptr(), ds.data_ptr(), dv.data_ptr(), cap, 0)
        rc = L.mz_run_device(ctx.handle, 0, C.byref(p), d_in.data_ptr(), 1, n, a, b, C.byref(out))
        assert rc == 0
        m = out.count
        got_p.append(dp[:m].cpu().numpy().view(np.uint32))
        got_s.append(ds[:m].cpu().numpy().view(np.uint32))
        got_v.append(dv[:m].cpu().numpy().view(np.uint64))
    assert np.array_equal(np.concatenate(got_p), epos)
    assert np.array_equal(np.concatenate(got_s), esk)
    assert np.array_equal(np.concatenate(got_v), evals)
    # capacity too small -> MZ_ERR_CAPACITY with the needed count
    dp = torch.empty(10, dtype=torch.int32, device="cuda")
    out = ffi.MzOut(dp.data_ptr(), dp.data_ptr(), dp.data_ptr(), 10, 0)
    p.want_sk, p.value_bits = 0, 0
    rc = L.mz_run_device(ctx.handle, 0, C.byref(p), d_in.data_ptr(), 1, n, 0, 0, C.byref(out))
    assert rc == 8 and out.count == len(epos)


def test_pipelined_host_path_chunk_seams(sm, oracle, monkeypatch):
    """mz_run's chunk-pipelined path (H2D / kernel / D2H overlap): force tiny chunks so many
    chunk seams fall inside the sequence; result must equal the single-shot run."""
    monkeypatch.setenv("MZ_PIPELINE_MIN_WINDOWS", "1")
    n = 400_000
    packed = oracle.synth_packed(17, n + 2)
    for chunk in ("777", "65536", "100000"):
        monkeypatch.setenv("MZ_CHUNK_WINDOWS", chunk)
        for (k, w, c, mode) in ((31, 19, True, 0), (21, 11, False, 0), (31, 11, True, 1), (9, 5, False, 2), (40, 40, False, 0)):
            _check_case(sm, oracle, packed, 2, n, k, w, c, mode)


def test_pipelined_host_path_default_chunks_pageable(sm, oracle, monkeypatch):
    """The default chunk of the one-device pipeline is 2^25 windows = 8 MiB + 12 bytes of packed
    input.  Pageable callers go through a bounce buffer filled by a multi-threaded copy whose
    slices once did not add up to the whole chunk (8 MiB / 16 threads is a multiple of the 4 KiB
    slice granule, the 12 odd bytes -- the bases of the chunk's last windows -- stayed behind)."""
    monkeypatch.setenv("MZ_PIPELINE_MIN_WINDOWS", "1")
    k, w = 31, 19
    n = (1 << 25) + (1 << 23) + 12345
    packed = oracle.synth_packed(9, n + 8)  # numpy array: pageable
    # chunk bytes just above a multiple of (threads x 4 KiB)
    for chunk_w in (1 << 25, (1 << 24) + 16, 12 * 4096 * 16 * 4 + 8):
        monkeypatch.setenv("MZ_CHUNK_WINDOWS", str(chunk_w))
        _check_case(sm, oracle, packed, 0, n, k, w, True, 0)


def _oracle_batch(oracle, packed, starts, lens, k, w, canonical, mode, want_sk):
    pr = oracle.make_params(k, w, canonical=canonical, mode=mode)
    offs, pos, sks, vals = [0], [], [], []
    length = k if mode == 0 else k + w - 1
    for s, n in zip(starts, lens):
        p, sk = oracle.run(packed, int(s), int(n), pr, want_sk=want_sk)
        pos.append(p)
        if want_sk:
            sks.append(sk)
        if length <= 32:
            vals.append(oracle.values_u64(packed, int(s), length, canonical, p))
        offs.append(offs[-1] + len(p))
    cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt)
    return (np.array(offs, dtype=np.uint64), cat(pos, np.uint32), cat(sks, np.uint32) if want_sk else None,
            cat(vals, np.uint64) if length <= 32 else None)


def test_batch_fixed_stride_short_reads(sm, oracle):
    """BASELINE config 5 shape: 150 bp reads at a 38-byte stride, canonical k=21 w=11; every
    read must give exactly what a per-read call gives (bench/src/bin/paper.rs:98-105)."""
    n_reads, read_len, stride = 5000, 150, 38
    packed = oracle.synth_packed(31, n_reads * stride * 4 + 64)
    starts = np.arange(n_reads, dtype=np.uint64) * (stride * 4)
    lens = np.full(n_reads, read_len, dtype=np.uint32)
    for (k, w, c, mode) in ((21, 11, True, 0), (21, 11, False, 0), (15, 10, False, 1), (31, 5, True, 1)):
        b = _builder(sm, k, w, c, mode)
        sk = sm.U32Vec()
        bb = b.super_kmers(sk) if mode == 0 else b
        offs, pos, sks, vals = bb.run_batch(packed, stride_bytes=stride, read_len=read_len, n_reads=n_reads)
        eo, ep, es, ev = _oracle_batch(oracle, packed, starts, lens, k, w, c, mode, mode == 0)
        assert np.array_equal(offs, eo) and np.array_equal(pos, ep), (k, w, c, mode)
        if mode == 0:
            assert np.array_equal(sks, es)
        if ev is not None:
            assert np.array_equal(vals, ev)


def test_batch_ragged_reads(sm, oracle):
    """Ragged batch: empty, too-short and unaligned reads, including w > 32 (generic kernel)."""
    rng = np.random.default_rng(8)
    n_reads = 700
    lens = rng.integers(0, 400, n_reads).astype(np.uint32)
    lens[:5] = [0, 1, 30, 31, 32]
    gaps = rng.integers(0, 9, n_reads)
    starts = np.zeros(n_reads, dtype=np.uint64)
    cur = 3
    for i in range(n_reads):
        cur += int(gaps[i])
        starts[i] = cur
        cur += int(lens[i])
    packed = oracle.synth_packed(77, cur + 64)
    for (k, w, c, mode) in ((21, 11, True, 0), (5, 4, False, 0), (9, 40, False, 0), (11, 33, True, 1), (7, 3, True, 2)):
        b = _builder(sm, k, w, c, mode)
        offs, pos, _, vals = b.run_batch(packed, starts=starts, lens=lens)
        eo, ep, _, ev = _oracle_batch(oracle, packed, starts, lens, k, w, c, mode, False)
        assert np.array_equal(offs, eo) and np.array_equal(pos, ep), (k, w, c, mode)
        if ev is not None:
            assert np.array_equal(vals, ev)


def test_cpp_mirror_header(sm, tmp_path):
    """include/simd_minimizers.hpp (the C++ host mirror) compiled with g++ against libmzb200.so
    reproduces the reference's doc-test vectors."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "simd-minimizers_b200")
    exe = str(tmp_path / "cpp_mirror_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp_mirror_test.cpp"), "-o", exe,
                           "-L", libdir, "-lmzb200", "-Wl,-rpath," + libdir])
    out = subprocess.check_output([exe]).decode()
    assert "cpp mirror ok" in out


def test_full_size_genome_properties(sm, oracle):
    """BASELINE config 2/3 at full size (3.1 Gbp, canonical k=31 w=19, pos + super-k-mer starts +
    u64 values): size-independent properties, exact oracle compare on the first / last million
    windows and random tiles, and checksum agreement between the chunk-pipelined host path and a
    single device-resident launch (two different seam decompositions of the same sequence)."""
    import ctypes as C
    import importlib

    import torch

    import bench

    n, k, w = 3_100_000_000, 31, 19
    l = k + w - 1
    nwin = n - l + 1
    packed, off = bench.synth_packed_range(bench.SEED, 0, n)
    packed = np.ascontiguousarray(packed)
    seq = sm.PackedSeq(packed, off, n)
    # positions, super-k-mer starts and values fused through the chunk pipeline ...
    p, s, vals = sm.canonical_minimizers(k, w).super_kmers(sm.U32Vec()).run_with_values(seq, 64)
    # ... and the same through the reference-shaped calls: positions only, lazy values (mz_values)
    pos, sk = sm.U32Vec(), sm.U32Vec()
    out = sm.canonical_minimizers(k, w).super_kmers(sk).run(seq, pos)
    assert np.array_equal(pos.array, p) and np.array_equal(sk.array, s)
    assert np.array_equal(out.values_u64(), vals)
    del pos, sk, out
    m = len(p)
    assert len(s) == m and len(vals) == m
    assert abs(m / n - 2.0 / (w + 1)) < 0.002                      # density of random minimizers
    assert int(p.max()) <= n - k and s[0] == 0 and int(s[-1]) < nwin
    assert (np.diff(s.astype(np.int64)) > 0).all()                 # one emission per window at most, in order
    assert (p[1:] != p[:-1]).all()                                 # adjacent duplicates removed
    d = p.astype(np.int64) - s.astype(np.int64)
    assert d.min() >= 0 and d.max() <= w - 1                       # minimizer lies inside its first window
    pr = oracle.make_params(k, w, canonical=True)
    rng = np.random.default_rng(1)
    ranges = [(0, 1_000_000), (nwin - 1_000_000, nwin)] + [
        (a, a + 200_000) for a in (int(x) for x in rng.integers(1_000_000, nwin - 2_000_000, 3))]
    for a, b in ranges:
        epos, esk = oracle.run_range(packed, off, n, pr, a, b, want_sk=True)
        i0, i1 = np.searchsorted(s, a, "left"), np.searchsorted(s, b, "left")
        assert np.array_equal(p[i0:i1], epos) and np.array_equal(s[i0:i1], esk), (a, b)
        assert np.array_equal(vals[i0:i1], oracle.values_u64(packed, off, k, True, epos)), (a, b)
    # same sequence, one device-resident launch: identical streams (checksums + exact on device)
    ffi = importlib.import_module("simd-minimizers_b200._ffi")
    L = ffi.lib()
    d_in = torch.from_numpy(packed).cuda()
    dp = torch.empty(m, dtype=torch.int32, device="cuda")
    ds = torch.empty(m, dtype=torch.int32, device="cuda")
    dv = torch.empty(m, dtype=torch.int64, device="cuda")
    prm = ffi.MzParams()
    L.mz_params_nthash(C.byref(prm), k, w, 0, 1)
    prm.want_sk, prm.value_bits = 1, 64
    o2 = ffi.MzOut(dp.data_ptr(), ds.data_ptr(), dv.data_ptr(), m, 0)
    ctx = sm.Context()
    assert L.mz_run_device(ctx.handle, 0, C.byref(prm), d_in.data_ptr(), off, n, 0, 0, C.byref(o2)) == 0
    assert o2.count == m
    hp = torch.from_numpy(p.view(np.int32)).cuda()
    assert torch.equal(hp, dp)
    del hp
    hs = torch.from_numpy(s.view(np.int32)).cuda()
    assert torch.equal(hs, ds)
    del hs
    hv = torch.from_numpy(vals.view(np.int64)).cuda()
    assert torch.equal(hv, dv)


def test_ascii_ingestion(sm, oracle):
    """AsciiSeq input (src/test.rs:55-110 asserts ASCII == packed): device-side packing must
    equal the host packing rule (c >> 1) & 3, and minimizers from ASCII equal those from packed."""
    rng = np.random.default_rng(4)
    for n in (0, 1, 15, 16, 17, 1000, 100_003):
        s = bytes(rng.choice(np.frombuffer(b"ACGTacgt", dtype=np.uint8), n))
        a = sm.AsciiSeq(s)
        packed = a.pack()
        ref = oracle.pack_ascii(s)
        assert np.array_equal(packed.data[:(n + 3) // 4], ref[:(n + 3) // 4])
        for (k, w, c) in ((5, 7, True), (21, 11, False)):
            b = (sm.canonical_minimizers if c else sm.minimizers)(k, w)
            pa, pp = sm.U32Vec(), sm.U32Vec()
            va = b.run(a, pa).values_u64()
            vp = b.run(packed, pp).values_u64()
            assert pa == pp and np.array_equal(va, vp)
            epos, _ = oracle.run(ref, 0, n, oracle.make_params(k, w, canonical=c))
            assert np.array_equal(pa.array, epos)


def test_very_large_w(sm, oracle):
    """w up to the reference's limit (w < 2^15, src/sliding_min.rs:92-95): the generic kernel keeps
    its two-stacks ring in global memory when it no longer fits shared memory."""
    packed = oracle.synth_packed(21, 200_000)
    for (k, w, c, n) in ((5, 801, True, 60_000), (8, 2000, False, 50_000), (3, 32767, True, 150_000),
                         (31, 1000, True, 20_000), (7, 4097, False, 4096 + 7 - 1)):
        if c and (k + w - 1) % 2 == 0:
            w += 1
        for mode in (0, 1):
            _check_case(sm, oracle, packed, 1, n, k, w, c, mode)


def test_batch_long_reads_are_cut_into_pieces(sm, oracle):
    """Reads longer than the per-thread record are processed as pieces of S windows; the result
    must still equal the per-read reference call (dedup across piece seams, CSR per read)."""
    rng = np.random.default_rng(10)
    lens = np.array([0, 20, 150, 289, 290, 318, 319, 320, 1000, 5000, 37, 12_345, 600, 49, 48], dtype=np.uint32)
    lens = np.concatenate([lens, rng.integers(0, 3000, 40).astype(np.uint32)])
    starts = np.zeros(len(lens), dtype=np.uint64)
    cur = 1
    for i, n in enumerate(lens):
        starts[i] = cur
        cur += int(n) + int(rng.integers(0, 5))
    packed = oracle.synth_packed(55, cur + 64)
    for (k, w, c, mode) in ((31, 19, True, 0), (21, 11, False, 0), (15, 9, True, 1), (9, 41, False, 0), (5, 3, False, 2)):
        b = _builder(sm, k, w, c, mode)
        sk = sm.U32Vec()
        bb = b.super_kmers(sk) if mode == 0 else b
        offs, pos, sks, vals = bb.run_batch(packed, starts=starts, lens=lens)
        eo, ep, es, ev = _oracle_batch(oracle, packed, starts, lens, k, w, c, mode, mode == 0)
        assert np.array_equal(offs, eo), (k, w, c, mode)
        assert np.array_equal(pos, ep), (k, w, c, mode)
        if mode == 0:
            assert np.array_equal(sks, es), (k, w, c, mode)
        if ev is not None:
            assert np.array_equal(vals, ev), (k, w, c, mode)


def test_degenerate_and_dense_outputs_at_scale(sm, oracle):
    """Worst cases for the ordered compaction at a size that spans many tiles and chunks:
    homopolymer (every window emits, leftmost == window start), short periods (ties in every
    window: the strand fix-up runs in every block), w=1 (density 1) and w=2.  The capacity
    estimate is exceeded on purpose, so the exact-size re-run path is exercised too."""
    n = 20_000_000
    pats = {"polyA": b"A", "polyG": b"G", "AC": b"AC", "ACGTT": b"ACGTT"}
    for name, pat in pats.items():
        s = (pat * (n // len(pat) + 1))[:n]
        packed = oracle.pack_ascii(s)
        for (k, w, c, mode) in ((31, 19, True, 0), (21, 11, False, 0), (31, 11, True, 1)):
            _check_case(sm, oracle, packed, 0, n, k, w, c, mode)
    rnd = oracle.synth_packed(9, n)
    for (k, w, c, mode) in ((5, 1, True, 0), (8, 2, False, 0), (31, 1, True, 1), (9, 3, True, 2)):
        _check_case(sm, oracle, rnd, 0, n, k, w, c, mode)


def test_random_differential_fuzz(sm, oracle):
    """Seeded random (k, w, len, offset, builder, hasher, mode) combinations against the oracle:
    broad coverage of kernel instances (every w <= 32 hits the W-specialised kernel, larger w the
    generic one) and of segment / tile / chunk boundaries."""
    rng = np.random.default_rng(20261017)
    base = oracle.synth_packed(4242, 70_000)
    for it in range(4000):
        k = int(rng.integers(1, 70))
        w = int(rng.integers(1, 34)) if rng.random() < 0.85 else int(rng.integers(34, 120))
        mode = int(rng.choice([0, 0, 0, 1, 2]))
        canonical = bool(rng.integers(0, 2))
        if canonical and (k + w - 1) % 2 == 0:
            w += 1
        if mode == 2 and w % 2 == 0:
            mode = 1
        r = rng.random()
        n = int(rng.integers(0, 300)) if r < 0.3 else (int(rng.integers(300, 12_000)) if r < 0.9 else int(rng.integers(12_000, 60_000)))
        off = int(rng.integers(0, 16))
        kind = "nt" if rng.random() < 0.7 else "mul"
        hash_canon = canonical or (rng.random() < 0.2)
        _check_case(sm, oracle, base, off, n, k, w, canonical, mode, kind=kind, hash_canon=hash_canon)


@pytest.mark.parametrize("chunk", ["16", "48", "1000"])
def test_batch_chunk_pipeline_seams(sm, oracle, monkeypatch, chunk):
    """mz_run_batch streams reads through the device in chunks (MZ_BATCH_CHUNK_READS forces tiny
    ones): fixed stride, ragged reads in storage order (with long reads cut into pieces), and
    reads that are NOT in storage order (single-chunk fallback) all give the per-read result,
    with CSR offsets rebased across chunks."""
    monkeypatch.setenv("MZ_BATCH_CHUNK_READS", chunk)
    rng = np.random.default_rng(21)
    # fixed stride
    n_reads, read_len, stride = 333, 150, 38
    packed = oracle.synth_packed(5, n_reads * stride * 4 + 64)
    starts = np.arange(n_reads, dtype=np.uint64) * (stride * 4)
    lens = np.full(n_reads, read_len, dtype=np.uint32)
    sk = sm.U32Vec()
    offs, pos, sks, vals = sm.canonical_minimizers(21, 11).super_kmers(sk).run_batch(
        packed, stride_bytes=stride, read_len=read_len, n_reads=n_reads)
    eo, ep, es, ev = _oracle_batch(oracle, packed, starts, lens, 21, 11, True, 0, True)
    assert np.array_equal(offs, eo) and np.array_equal(pos, ep) and np.array_equal(sks, es)
    assert np.array_equal(vals, ev)
    # ragged, storage order, some long reads
    n_reads = 260
    lens = rng.integers(0, 500, n_reads).astype(np.uint32)
    lens[[7, 100, 259]] = [4000, 9000, 700]
    starts = np.zeros(n_reads, dtype=np.uint64)
    cur = 2
    for i in range(n_reads):
        starts[i] = cur
        cur += int(lens[i]) + int(rng.integers(0, 7))
    packed = oracle.synth_packed(6, cur + 64)
    for (k, w, c, mode) in ((31, 19, True, 0), (15, 10, False, 1), (9, 40, False, 0)):
        b = _builder(sm, k, w, c, mode)
        offs, pos, _, vals = b.run_batch(packed, starts=starts, lens=lens)
        eo, ep, _, ev = _oracle_batch(oracle, packed, starts, lens, k, w, c, mode, False)
        assert np.array_equal(offs, eo) and np.array_equal(pos, ep), (k, w, c, mode)
        if ev is not None:
            assert np.array_equal(vals, ev)
    # shuffled order: cannot be streamed
    perm = rng.permutation(n_reads)
    offs, pos, _, vals = sm.canonical_minimizers(21, 11).run_batch(packed, starts=starts[perm], lens=lens[perm])
    eo, ep, _, ev = _oracle_batch(oracle, packed, starts[perm], lens[perm], 21, 11, True, 0, False)
    assert np.array_equal(offs, eo) and np.array_equal(pos, ep) and np.array_equal(vals, ev)


def test_long_windows_subwindow_path(sm, oracle):
    """32 < w <= 255 runs on the W-specialised kernel as the minimum over shifted sub-windows
    (one to ten ring taps); w >= 256 on the runtime-w kernel.  Window lengths on both sides of
    every tap-count boundary, all builders, random and all-ties (homopolymer) input."""
    rng = np.random.default_rng(77)
    n = 60_000
    packed = oracle.synth_packed(99, n + 8)
    homo = np.zeros(n // 4 + 16, dtype=np.uint8)  # AAAA...: every window is one big tie
    mixed = packed.copy()
    mixed[2000:5000] = 0x55                       # a long CCCC... run inside random sequence
    ws = [33, 34, 47, 48, 49, 50, 64, 65, 71, 72, 73, 96, 97, 120, 121, 128, 200, 254, 255, 256, 300]
    for w in ws:
        for canonical in (False, True):
            k = int(rng.integers(1, 40))
            if canonical and (k + w - 1) % 2 == 0:
                k += 1
            for mode in (0, 1, 2):
                if mode == 2 and w % 2 == 0:
                    continue
                for data, off, nn in ((packed, 3, n), (mixed, 0, 30_011), (homo, 1, 9_000 + w)):
                    _check_case(sm, oracle, data, off, nn, k, w, canonical, mode,
                                kind="mul" if w % 3 == 0 else "nt")
    # forward builder with a canonical hasher, and short inputs around l
    for w in (33, 64, 255):
        for nn in (w + 4, w + 5, 2 * w + 9, 5 * w):
            _check_case(sm, oracle, packed, 2, nn, 5, w, False, 0, hash_canon=True)
            _check_case(sm, oracle, packed, 2, nn, 5 if w % 2 else 6, w, True, 0)


def test_very_large_k(sm, oracle):
    """k far beyond the value widths (hash warm-up of thousands of bases, halo several times the
    segment), combined with short, long and runtime-w windows; positions only."""
    n = 300_000
    packed = oracle.synth_packed(3, n + 8)
    for (k, w, canonical, mode) in ((501, 19, True, 0), (2001, 11, True, 0), (1000, 50, False, 0),
                                    (777, 33, True, 1), (4097, 5, False, 2), (300, 300, False, 0),
                                    (301, 301, True, 0)):
        b = _builder(sm, k, w, canonical, mode)
        pr = oracle.make_params(k, w, canonical=canonical, mode=mode)
        for off, nn in ((1, n), (0, k + w + 5), (2, 3 * (k + w))):
            want, _ = oracle.run(packed, off, nn, pr, "stream")
            got = b.run_once(sm.PackedSeq(packed, off, nn))
            assert np.array_equal(got, want), (k, w, canonical, mode, off, nn)


def test_concurrent_host_threads(sm, oracle):
    """The reference is called from many rayon workers at once (bench/src/bin/paper.rs:442-459);
    here every host thread owns an mz_ctx.  Four threads, different parameters, concurrent calls
    (ctypes releases the GIL), results must equal the single-threaded ones."""
    import threading

    n = 3_000_000
    packed = oracle.synth_packed(17, n + 8)
    cases = [(31, 19, True, 0), (21, 11, False, 0), (15, 41, True, 1), (9, 5, False, 2)]
    expect = []
    for (k, w, c, mode) in cases:
        pr = oracle.make_params(k, w, canonical=c, mode=mode)
        expect.append(oracle.run(packed, 1, n, pr, "stream")[0])
    errors = []

    def worker(i):
        try:
            k, w, c, mode = cases[i]
            ctx = sm.Context()
            b = _builder(sm, k, w, c, mode).context(ctx)
            for _ in range(5):
                got = b.run_once(sm.PackedSeq(packed, 1, n))
                if not np.array_equal(got, expect[i]):
                    errors.append((i, "mismatch"))
            ctx.close()
        except Exception as e:  # noqa: BLE001
            errors.append((i, repr(e)))

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(len(cases))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


def test_c_example_runs(tmp_path):
    """examples/minimal.c (plain C99 caller of the ABI) reproduces the README vector."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "simd-minimizers_b200")
    exe = str(tmp_path / "minimal")
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(root, "include"),
                           os.path.join(root, "examples", "minimal.c"), "-L", lib, "-lmzb200",
                           "-Wl,-rpath," + lib, "-o", exe])
    out = subprocess.check_output([exe]).decode()
    assert "pos 15 value 817" in out


def test_pipelined_transfer_codec(sm, oracle, monkeypatch):
    """Positions / super-k-mer starts cross PCIe as int8 deltas in the chunk-pipelined mz_run
    (minimizers, w <= 127).  Inputs that stress the codec: ties everywhere (homopolymer and
    dinucleotide repeats: the strand rule steps backwards), the largest coded window (w = 127,
    deltas up to +-127), the first window length that is copied verbatim (w = 128), entry counts
    around the 256-entry block size, and the codec switched off gives the same arrays."""
    monkeypatch.setenv("MZ_PIPELINE_MIN_WINDOWS", "1")
    n = 300_000
    rnd = oracle.synth_packed(23, n + 8)
    homo = np.zeros(n // 4 + 16, dtype=np.uint8)                 # AAAA...
    dinuc = np.full(n // 4 + 16, 0x11 * 0 + 0b01000100, dtype=np.uint8)  # ACAC...
    tga = np.full(n // 4 + 16, 0b10110010, dtype=np.uint8)       # T A G T repeats: TG-rich / poor windows
    mixed = rnd.copy()
    mixed[1000:9000] = 0b11101110                                # long GTGT run inside random sequence
    for chunk in ("1000", "70001"):
        monkeypatch.setenv("MZ_CHUNK_WINDOWS", chunk)
        for data in (rnd, homo, dinuc, tga, mixed):
            for (k, w, c) in ((31, 19, True), (5, 127, True), (4, 127, False), (5, 128, False), (6, 3, False), (1, 1, True)):
                if c and (k + w - 1) % 2 == 0:
                    k += 1
                _check_case(sm, oracle, data, 1, n, k, w, c, 0)
    # tiny outputs: 0, 1, 255, 256, 257 entries per chunk
    for nn in (60, 61, 300, 2600, 2700):
        monkeypatch.setenv("MZ_CHUNK_WINDOWS", "100000")
        _check_case(sm, oracle, rnd, 0, nn, 31, 19, True, 0)
    # codec off == codec on
    seq = sm.PackedSeq(mixed, 2, n)
    a = sm.canonical_minimizers(21, 11).run_once(seq)
    monkeypatch.setenv("MZ_NO_POS_DELTA", "1")
    b = sm.canonical_minimizers(21, 11).run_once(seq)
    assert np.array_equal(a, b)
