"""CPU tests: the oracle against every known-answer vector the reference holds for this path
(SURVEY.md section 8c), plus naive == streaming differential checks and the rc-symmetry
properties of src/test.rs:113-152, 642-708."""
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")


@pytest.fixture(scope="module")
def vec():
    with open(GOLDEN) as f:
        return json.load(f)


def test_doc_forward(oracle, vec):
    v = vec["doc_forward"]                      # src/lib.rs:92-99
    s = v["seq"].encode()
    p = oracle.pack_ascii(s)
    for algo in ("naive", "stream"):
        pos, _ = oracle.run(p, 0, len(s), oracle.make_params(v["k"], v["w"], canonical=False), algo)
        assert pos.tolist() == v["pos"]


def test_doc_canonical_and_values(oracle, vec):
    v = vec["doc_canonical"]                    # src/lib.rs:109-129, README.md:41-50, src/test.rs:406-415
    s = v["seq"].encode()
    p = oracle.pack_ascii(s)
    pr = oracle.make_params(v["k"], v["w"], canonical=True)
    for algo in ("naive", "stream"):
        pos, _ = oracle.run(p, 0, len(s), pr, algo)
        assert pos.tolist() == v["pos"]
    vals = oracle.values_u64(p, 0, v["k"], True, pos)
    assert vals.tolist() == v["values"]
    rc = oracle.revcomp(p, 0, len(s))           # src/lib.rs:132-140
    rpos, _ = oracle.run(rc, 0, len(s), pr)
    assert rpos.tolist() == v["rc_pos"]
    rvals = oracle.values_u64(rc, 0, v["k"], True, rpos)
    assert rvals.tolist()[::-1] == v["values"]


def test_collect_and_dedup_scalar(oracle, vec):
    for case in vec["collect_dedup"]:           # src/test.rs:335-356
        pos, sk = oracle.collect_dedup(np.array(case["in"], dtype=np.uint32))
        assert pos.tolist() == case["out"]
        assert sk.tolist() == case["idx"]


def test_syncmer_predicates(oracle, vec):
    for case in vec["syncmers_scalar"]:         # src/test.rs:485-515
        out = oracle.collect_syncmers(np.array(case["min_pos"], dtype=np.uint32), case["w"], case["open"])
        assert out.tolist() == case["out"]


def test_closed_syncmer_values_all_g(oracle):
    n = 100                                     # src/test.rs:578-597
    p = oracle.pack_ascii(b"G" * n)
    for k in range(1, 10):
        for w in range(1, 10):
            pr = oracle.make_params(k, w, canonical=False, mode=oracle.CLOSED_SYNCMER)
            pos, _ = oracle.run(p, 0, n, pr)
            l = k + w - 1
            assert len(pos) == n - l + 1
            vals = oracle.values_u64(p, 0, l, False, pos)
            assert (vals == (1 << (2 * l)) - 1).all()


def _grid(rng, count):
    ks = [1, 2, 3, 4, 5, 31, 32, 33, 63, 64, 65]
    ws = [1, 2, 3, 4, 5, 31, 32, 33, 63, 64, 65]
    for _ in range(count):
        k = int(rng.choice(ks)) if rng.random() < 0.6 else int(rng.integers(6, 100))
        w = int(rng.choice(ws)) if rng.random() < 0.6 else int(rng.integers(6, 100))
        n = int(rng.integers(0, 100)) if rng.random() < 0.6 else int(rng.integers(100, 1500))
        yield k, w, n, int(rng.integers(0, 4))


def test_naive_equals_stream(oracle):
    """src/test.rs:55-110 (naive == scalar), on the reference's (k, w, len, offset) grid."""
    rng = np.random.default_rng(2024)
    for k, w, n, off in _grid(rng, 400):
        seq = oracle.synth_packed(int(rng.integers(0, 1 << 40)), n + off)
        for kind in ("nt", "mul"):
            for canon in (False, True):
                if canon and (k + w - 1) % 2 == 0:
                    continue
                for mode in (0, 1, 2):
                    if mode == 2 and w % 2 == 0:
                        continue
                    pr = oracle.make_params(k, w, canonical=canon, mode=mode,
                                            hasher=oracle.make_hasher(kind, canon))
                    a = oracle.run(seq, off, n, pr, "naive", want_sk=(mode == 0))
                    b = oracle.run(seq, off, n, pr, "stream", want_sk=(mode == 0))
                    assert np.array_equal(a[0], b[0]), (k, w, n, off, kind, canon, mode)
                    if mode == 0:
                        assert np.array_equal(a[1], b[1])


def test_rc_symmetry(oracle):
    """src/test.rs:113-152 (minimizers) and 642-708 (syncmers)."""
    rng = np.random.default_rng(7)
    done = 0
    for k, w, n, off in _grid(rng, 600):
        l = k + w - 1
        if l % 2 == 0 or n < l:
            continue
        seq = oracle.synth_packed(int(rng.integers(0, 1 << 40)), n + off)
        rc = oracle.revcomp(seq, off, n)
        for mode in (0, 1, 2):
            if mode == 2 and w % 2 == 0:
                continue
            pr = oracle.make_params(k, w, canonical=True, mode=mode)
            fpos, _ = oracle.run(seq, off, n, pr)
            rpos, _ = oracle.run(rc, 0, n, pr)
            length = k if mode == 0 else l
            assert len(fpos) == len(rpos)
            assert (fpos.astype(np.int64) + rpos[::-1].astype(np.int64) == n - length).all()
            if length <= 32:
                fv = oracle.values_u64(seq, off, length, True, fpos)
                rv = oracle.values_u64(rc, 0, length, True, rpos)
                assert np.array_equal(fv, rv[::-1])
            elif length <= 64:
                fv = oracle.values_u128(seq, off, length, True, fpos)
                rv = oracle.values_u128(rc, 0, length, True, rpos)
                assert np.array_equal(fv, rv[::-1])
        done += 1
    assert done > 50


def test_run_range_and_mt_match_single(oracle):
    rng = np.random.default_rng(3)
    n = 20000
    seq = oracle.synth_packed(5, n + 2)
    for (k, w, canon, mode) in ((31, 19, True, 0), (21, 11, False, 0), (5, 7, True, 1), (9, 5, False, 2)):
        pr = oracle.make_params(k, w, canonical=canon, mode=mode)
        pos, sk = oracle.run(seq, 2, n, pr, want_sk=(mode == 0))
        mpos, msk, mval = oracle.run_mt(seq, 2, n, pr, threads=5, want_sk=(mode == 0), want_val=(k <= 32 and mode == 0))
        assert np.array_equal(pos, mpos)
        if mode == 0:
            assert np.array_equal(sk, msk)
            assert np.array_equal(mval, oracle.values_u64(seq, 2, k, canon, pos))
        nwin = n - (k + w - 1) + 1
        cuts = sorted(set([0, nwin] + [int(x) for x in rng.integers(0, nwin, 4)]))
        parts = [oracle.run_range(seq, 2, n, pr, a, b)[0] for a, b in zip(cuts[:-1], cuts[1:])]
        assert np.array_equal(np.concatenate(parts), pos)


def test_invalid_params(oracle):
    seq = oracle.synth_packed(1, 100)
    with pytest.raises(ValueError):
        oracle.run(seq, 0, 100, oracle.make_params(4, 3, canonical=True))       # even l
    with pytest.raises(ValueError):
        oracle.run(seq, 0, 100, oracle.make_params(5, 4, canonical=False, mode=oracle.OPEN_SYNCMER))


def test_collect_skip_max_vectors(oracle):
    """src/test.rs:359-399: SKIPPED elements vanish; the comparison is against the stream
    element just before (each SIMD lane carries the same stream, so lane 0's output is a
    prefix of the reference's `out`)."""
    x = oracle.SKIPPED
    got = oracle.collect_dedup_skip_max(np.array([0, 1, 1, x, 2, 3, x, x, 4], dtype=np.uint32))
    assert got.tolist() == [0, 1, 2, 3, 4]
    got = oracle.collect_dedup_skip_max(np.array([1, x, x, x, x, x, x, 2, x, x, x, x], dtype=np.uint32))
    assert got.tolist() == [1, 2]
    # without SKIP_MAX the same stream keeps one x per run (first half of the reference test)
    pos, _ = oracle.collect_dedup(np.array([0, 1, 1, x, 2, 3, x, x, 4], dtype=np.uint32))
    assert pos.tolist() == [0, 1, x, 2, 3, x, 4]


def _random_n_ascii(rng, n, frac, runs):
    s = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
    for _ in range(int(n * frac)):
        s[rng.integers(0, n)] = ord("N")
    for _ in range(runs):
        a = int(rng.integers(0, n))
        s[a:a + int(rng.integers(1, 200))] = ord("N")
    return s.tobytes()


def test_skip_ambiguous_properties(oracle):
    """src/test.rs:429-482: no SKIPPED in the output and no ambiguous base inside any reported
    k-mer; plus naive == streaming, and the equivalent 'filter' formulation: an element of the
    plain run survives iff its run of windows contains a window without ambiguous bases."""
    rng = np.random.default_rng(5)
    for it in range(60):
        n = int(rng.integers(1, 600))
        ascii_ = _random_n_ascii(rng, n, 0.01 if it % 2 else 0.0, it % 4)
        packed, amb = oracle.pack_ascii_n(ascii_)
        ambits = np.unpackbits(amb, bitorder="little")[:n]
        assert np.array_equal(ambits, np.frombuffer(ascii_, dtype=np.uint8) == ord("N"))
        k = int(rng.integers(1, 40))
        w = int(rng.integers(1, 40))
        if (k + w - 1) % 2 == 0:
            w += 1
        for mode in (oracle.MINIMIZER, oracle.CLOSED_SYNCMER, oracle.OPEN_SYNCMER):
            if mode == oracle.OPEN_SYNCMER and w % 2 == 0:
                continue
            pr = oracle.make_params(k, w, canonical=True, mode=mode)
            a = oracle.run_skip_ambiguous(packed, 0, n, amb, 0, pr, "naive")
            b = oracle.run_skip_ambiguous(packed, 0, n, amb, 0, pr, "stream")
            assert np.array_equal(a, b)
            l = k + w - 1
            span = k if mode == oracle.MINIMIZER else l
            for p in a.tolist():
                assert p != oracle.SKIPPED
                assert not ambits[p:p + span].any()
            # filter formulation
            nwin = max(0, n - l + 1)
            clean = np.array([not ambits[j:j + l].any() for j in range(nwin)], dtype=bool)
            if mode == oracle.MINIMIZER:
                pos, sk = oracle.run(packed, 0, n, pr, "stream", want_sk=True)
                ends = np.append(sk[1:], nwin)
                keep = [bool(clean[s:e].any()) for s, e in zip(sk.tolist(), ends.tolist())]
                assert np.array_equal(pos[np.array(keep, dtype=bool)] if len(pos) else pos, a)
            else:
                pos, _ = oracle.run(packed, 0, n, pr, "stream")
                assert np.array_equal(pos[clean[pos]] if len(pos) else pos, a)
    # offsets into both streams
    ascii_ = _random_n_ascii(rng, 500, 0.01, 2)
    packed, amb = oracle.pack_ascii_n(ascii_)
    for off in (1, 2, 3, 5, 9):
        p2, a2 = oracle.pack_ascii_n(ascii_[off:])
        pr = oracle.make_params(7, 5, canonical=True)
        assert np.array_equal(oracle.run_skip_ambiguous(packed, off, 500 - off, amb, off, pr),
                              oracle.run_skip_ambiguous(p2, 0, 500 - off, a2, 0, pr))


def test_position_delta_bound(oracle):
    """The PCIe transfer codec of mz_run stores consecutive positions / super-k-mer starts as int8
    deltas for w <= 127.  The bound it relies on, checked on the oracle: p[i+1] - p[i] lies in
    [-(w-2), w] (the strand rule may step back inside a tie) and sk[i+1] - sk[i] in [1, w]."""
    rng = np.random.default_rng(9)
    n = 20_000
    seqs = [oracle.synth_packed(4, n + 8), np.zeros(n // 4 + 16, dtype=np.uint8),
            np.full(n // 4 + 16, 0b01000100, dtype=np.uint8), np.full(n // 4 + 16, 0b10110010, dtype=np.uint8)]
    mixed = seqs[0].copy()
    mixed[500:1500] = 0b11101110
    seqs.append(mixed)
    saw_backward = False
    for packed in seqs:
        for _ in range(12):
            k = int(rng.integers(1, 34))
            w = int(rng.integers(1, 128))
            canonical = bool(rng.integers(0, 2))
            if canonical and (k + w - 1) % 2 == 0:
                w = w + 1 if w < 127 else w - 1
            pr = oracle.make_params(k, w, canonical=canonical)
            pos, sk = oracle.run(packed, 1, n, pr, "stream", want_sk=True)
            if len(pos) < 2:
                continue
            dp = np.diff(pos.astype(np.int64))
            ds = np.diff(sk.astype(np.int64))
            assert dp.max() <= w and dp.min() >= -(max(w, 2) - 2), (k, w, canonical, dp.min(), dp.max())
            assert ds.min() >= 1 and ds.max() <= w, (k, w, canonical, ds.min(), ds.max())
            saw_backward |= bool((dp < 0).any())
    assert saw_backward  # the tie-heavy inputs do exercise the negative range
