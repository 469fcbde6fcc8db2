"""Independent cross-check of oracle/mzoracle.c: a pure-Python restatement of the reference
semantics written from SURVEY.md Appendix A (ASCII in, no packing, no rolling state, no shared code
with the C oracle), compared on small random and tie-heavy inputs for every builder.
Reference citations: hash src/minimizers.rs:44-61 + seq-hash 0.2.0 (NtHasher constants pinned by
src/lib.rs:92-135), window argmin src/sliding_min.rs:102-129,190-195, strand rule
src/canonical.rs:19-29, dedup / super-k-mer index src/collect.rs:39-76, syncmers
src/syncmers.rs:33-37, values src/lib.rs:598-629, skip-ambiguous src/minimizers.rs:169-214."""
import numpy as np
import pytest

F = {"A": 0x95C60474, "C": 0x62A02B4C, "T": 0x82572324, "G": 0x4BE24456}
COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}
CODE = {"A": 0, "C": 1, "T": 2, "G": 3}
M32 = 0xFFFFFFFF


def rotl(x, r):
    r %= 32
    return ((x << r) | (x >> (32 - r))) & M32 if r else x


def py_hash(kmer, canonical, mul=False):
    k = len(kmer)
    tab = (lambda b: (CODE[b] * 0x27220A95) & M32) if mul else (lambda b: F[b])
    fw = rc = 0
    for j, b in enumerate(kmer):
        fw ^= rotl(tab(b), 7 * (k - 1 - j))
        rc ^= rotl(tab(COMP[b]), 7 * j)
    return (fw + rc) & M32 if canonical else fw


def py_run(seq, k, w, canonical, mode, mul=False, amb=None):
    n, l = len(seq), k + w - 1
    if canonical:
        assert l % 2 == 1
    keys = [py_hash(seq[i:i + k], canonical, mul) >> 16 for i in range(max(0, n - k + 1))]
    sel = []
    for j in range(max(0, n - l + 1)):
        win = keys[j:j + w]
        m = min(win)
        left = j + win.index(m)
        right = j + (w - 1 - win[::-1].index(m))
        if canonical:
            tg = sum(1 for b in seq[j:j + l] if b in "TG")
            p = left if 2 * tg > l else right
        else:
            p = left
        if amb is not None and any(amb[j:j + l]):
            p = None  # SKIPPED
        sel.append(p)
    pos, sk = [], []
    for j, p in enumerate(sel):
        if p is None:
            continue
        if mode == 0:
            if j == 0 or sel[j - 1] != p:  # the element before it in the stream, skipped or not
                pos.append(p)
                sk.append(j)
        elif mode == 1:
            if p == j or p == j + w - 1:
                pos.append(j)
        else:
            if p == j + w // 2:
                pos.append(j)
    return pos, sk


def py_value(seq, p, length, canonical):
    def enc(s):
        return sum(CODE[b] << (2 * i) for i, b in enumerate(s))
    s = seq[p:p + length]
    v = enc(s)
    if canonical:
        v = min(v, enc("".join(COMP[b] for b in reversed(s))))
    return v


def _cases(rng, count):
    out = []
    while len(out) < count:
        k, w = int(rng.integers(1, 12)), int(rng.integers(1, 12))
        canonical = bool(rng.integers(0, 2))
        if canonical and (k + w - 1) % 2 == 0:
            continue
        out.append((k, w, canonical))
    return out


def test_c_oracle_equals_python_restatement(oracle):
    rng = np.random.default_rng(1234)
    alphabets = ["ACGT", "AC", "A", "TG", "ACGTTTTTTT"]
    for it, (k, w, canonical) in enumerate(_cases(rng, 120)):
        n = int(rng.integers(0, 90))
        seq = "".join(rng.choice(list(alphabets[it % len(alphabets)]), size=n)) if n else ""
        packed = oracle.pack_ascii(seq.encode())
        for mul in (False, True):
            h = oracle.make_hasher("mul" if mul else "nt", canonical)
            for mode in (0, 1, 2):
                if mode == 2 and w % 2 == 0:
                    continue
                pr = oracle.make_params(k, w, canonical=canonical, mode=mode, hasher=h)
                want_pos, want_sk = py_run(seq, k, w, canonical, mode, mul)
                for algo in ("naive", "stream"):
                    pos, sk = oracle.run(packed, 0, n, pr, algo, want_sk=(mode == 0))
                    assert pos.tolist() == want_pos, (seq, k, w, canonical, mode, mul, algo)
                    if mode == 0:
                        assert sk.tolist() == want_sk, (seq, k, w, canonical, mode, mul, algo)
                length = k if mode == 0 else k + w - 1
                if want_pos and length <= 32:
                    vals = oracle.values_u64(packed, 0, length, canonical, np.array(want_pos, dtype=np.uint32))
                    assert vals.tolist() == [py_value(seq, p, length, canonical) for p in want_pos]


def test_c_oracle_skip_ambiguous_equals_python(oracle):
    rng = np.random.default_rng(99)
    for it, (k, w, _) in enumerate(_cases(rng, 80)):
        if (k + w - 1) % 2 == 0:
            w += 1
        n = int(rng.integers(0, 120))
        chars = rng.choice(list("ACGT"), size=n) if n else np.array([], dtype="<U1")
        for _ in range(n // 15):
            a = int(rng.integers(0, n))
            chars[a:a + int(rng.integers(1, 4))] = "N"
        seq = "".join(chars)
        packed, amb = oracle.pack_ascii_n(seq.encode())
        ambits = [c == "N" for c in seq]
        clean_seq = seq.replace("N", "G")  # (c >> 1) & 3 of 'N' is 3 = G; never looked at anyway
        for mode in (0, 1, 2):
            if mode == 2 and w % 2 == 0:
                continue
            pr = oracle.make_params(k, w, canonical=True, mode=mode)
            want, _ = py_run(clean_seq, k, w, True, mode, amb=ambits)
            for algo in ("naive", "stream"):
                got = oracle.run_skip_ambiguous(packed, 0, n, amb, 0, pr, algo)
                assert got.tolist() == want, (seq, k, w, mode, algo)


def test_python_restatement_reproduces_reference_vectors():
    """The Python restatement itself is pinned to the reference's known-answer vectors
    (src/lib.rs:92-99, 109-129, 132-135)."""
    assert py_run("ACGTGCTCAGAGACTCAG", 5, 7, False, 0)[0] == [4, 5, 8, 13]
    seq = "ACGTGCTCAGAGACTCAGAGGA"
    pos, _ = py_run(seq, 5, 7, True, 0)
    assert pos == [0, 7, 9, 15]
    assert [py_value(seq, p, 5, True) for p in pos] == [0b1011010001, 0b1100110001, 0b0100110011, 0b1100110001]
    rc = "".join(COMP[b] for b in reversed(seq))
    assert py_run(rc, 5, 7, True, 0)[0] == [2, 8, 10, 17]
