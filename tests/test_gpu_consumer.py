"""On-device consumer: mz_run_bucket_stats shards super-k-mers
(bench/src/minimizer.rs:3-36, Problem C) by their minimizer without moving the minimizer run over
PCIe.  Checked against histograms built on the host from the oracle's .super_kmers() run."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix64(x):
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def bucket_of(vals, nb):
    """floor(mix64(v) * nb / 2^64) with 64-bit arithmetic (nb < 2^32)."""
    h = mix64(vals.astype(np.uint64))
    with np.errstate(over="ignore"):
        hi, lo = h >> np.uint64(32), h & np.uint64(0xFFFFFFFF)
        return ((hi * np.uint64(nb) + ((lo * np.uint64(nb)) >> np.uint64(32))) >> np.uint64(32)).astype(np.int64)


def expected(oracle, packed, off, n, k, w, canonical, kind, nb):
    pr = oracle.make_params(k, w, canonical=canonical, hasher=oracle.make_hasher(kind, canonical))
    pos, sk = oracle.run(packed, off, n, pr, "stream", want_sk=True)
    vals = oracle.values_u64(packed, off, k, canonical, pos)
    nwin = max(0, n - (k + w - 1) + 1)
    b = bucket_of(vals, nb)
    lens = np.diff(np.concatenate([sk.astype(np.int64), [nwin]])) if len(sk) else np.zeros(0, dtype=np.int64)
    cnt = np.bincount(b, minlength=nb).astype(np.uint64)
    win = np.bincount(b, weights=lens, minlength=nb).astype(np.uint64)
    return cnt, win, len(pos), nwin


def test_bucket_stats_match_host_histograms(sm, oracle, monkeypatch):
    n = 2_000_003
    packed = oracle.synth_packed(77, n + 8)
    for (k, w, canonical, kind, nb, off) in ((31, 19, True, "nt", 1024, 3), (21, 11, False, "nt", 7, 0),
                                             (31, 19, True, "mul", 16384, 1), (5, 3, True, "nt", 1, 2), (32, 4, False, "nt", 4096, 0)):
        H = sm.NtHasher if kind == "nt" else sm.MulHasher
        b = (sm.canonical_minimizers if canonical else sm.minimizers)(k, w).hasher(H(k, canonical))
        ecnt, ewin, em, nwin = expected(oracle, packed, off, n, k, w, canonical, kind, nb)
        for chunk in (None, "100003", "7777"):
            if chunk:
                monkeypatch.setenv("MZ_CHUNK_WINDOWS", chunk)
            else:
                monkeypatch.delenv("MZ_CHUNK_WINDOWS", raising=False)
            for devs in (None, [0, 0, 0]):
                bb = b.context(sm.Context(devs)) if devs else b
                cnt, win, m = bb.bucket_stats(sm.PackedSeq(packed, off, n), nb)
                tag = (k, w, canonical, kind, nb, chunk, devs)
                assert m == em, tag
                assert np.array_equal(cnt, ecnt), tag
                assert np.array_equal(win, ewin), tag
                assert int(win.sum()) == nwin and int(cnt.sum()) == em
    monkeypatch.delenv("MZ_CHUNK_WINDOWS", raising=False)


def test_bucket_stats_edges(sm, oracle):
    packed = oracle.synth_packed(5, 4096)
    b = sm.canonical_minimizers(31, 19)
    # too short: nothing
    cnt, win, m = b.bucket_stats(sm.PackedSeq(packed, 0, 30), 16)
    assert m == 0 and cnt.sum() == 0 and win.sum() == 0
    # exactly one window
    cnt, win, m = b.bucket_stats(sm.PackedSeq(packed, 0, 49), 16)
    assert m == 1 and cnt.sum() == 1 and win.sum() == 1
    # homopolymer: every window its own super-k-mer, one bucket
    homo = np.zeros(4096, dtype=np.uint8)
    cnt, win, m = sm.minimizers(5, 4).bucket_stats(sm.PackedSeq(homo, 0, 1000), 64)
    assert m == 993 and (cnt > 0).sum() == 1 and int(win.sum()) == 993
    # errors: syncmers have no super-k-mers, k > 32 has no u64 value, too many buckets
    with pytest.raises(TypeError):
        sm.closed_syncmers(5, 3).bucket_stats(sm.PackedSeq(packed, 0, 100), 4)
    with pytest.raises(AssertionError):
        sm.minimizers(33, 3).bucket_stats(sm.PackedSeq(packed, 0, 100), 4)
    with pytest.raises(sm.MzError):
        b.bucket_stats(sm.PackedSeq(packed, 0, 100), 16385)


def test_bucket_stats_saves_the_bus(sm, oracle):
    """At 100 Mbp the consumer call must beat moving the run to the host (same context, pinned or not)."""
    import time

    n = 100_000_000
    packed = oracle.synth_packed(9, n + 8)
    seq = sm.PackedSeq(packed, 0, n)
    b = sm.canonical_minimizers(31, 19)
    b.bucket_stats(seq, 4096)
    t0 = time.perf_counter()
    cnt, win, m = b.bucket_stats(seq, 4096)
    t_dev = time.perf_counter() - t0
    t0 = time.perf_counter()
    pos, sk, vals = b.super_kmers(sm.U32Vec()).run_with_values(seq, 64)
    t_host = time.perf_counter() - t0
    assert m == len(pos)
    bk = bucket_of(vals, 4096)
    assert np.array_equal(cnt, np.bincount(bk, minlength=4096).astype(np.uint64))
    assert t_dev < t_host, (t_dev, t_host)
