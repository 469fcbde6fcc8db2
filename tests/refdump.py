"""Checker for tests/golden/reference_dump.json (written by tools/dump_reference_vectors.rs run
inside the reference crate).  Shared by tests/test_reference_dump.py (real dump, when present) and
its self-test (a dump synthesised from the oracle, so that this checker is itself exercised)."""
from __future__ import annotations

import numpy as np

CODE = {ord("A"): 0, ord("C"): 1, ord("T"): 2, ord("G"): 3, ord("a"): 0, ord("c"): 1, ord("t"): 2, ord("g"): 3}
M32 = 0xFFFFFFFF


def rotl(x, r):
    r &= 31
    return ((x << r) | (x >> (32 - r))) & M32 if r else x & M32


def lcg_dna(n: int, x: int = 42) -> bytes:
    """Same generator as tools/dump_reference_vectors.rs::lcg_dna."""
    out = bytearray()
    for _ in range(n):
        x = (x * 6364136223846793005 + 1442695040888963407) & (2**64 - 1)
        out.append(b"ACGT"[(x >> 33) & 3])
    return bytes(out)


def table_hash(text: bytes, i: int, k: int, f, c, rot: int, canonical: bool) -> int:
    """fw = XOR rotl(f[b], rot*(k-1-j)), rc = XOR rotl(c[b], rot*j), h = canonical ? fw + rc : fw
    (include/mz_b200.h)."""
    fw = rc = 0
    for j in range(k):
        b = CODE[text[i + j]]
        fw ^= rotl(f[b], rot * (k - 1 - j))
        rc ^= rotl(c[b], rot * j)
    return (fw + rc) & M32 if canonical else fw


def derive_tables(hash_rows: list[dict]) -> dict:
    """Per (hasher, seed): the per-base tables f, c and the rotation, read off the k = 1 and k = 2
    hashes, then verified against every dumped k.  Returns {(hasher, seed): dict(f, c, rot, ok, why)}."""
    groups: dict = {}
    for r in hash_rows:
        groups.setdefault((r["hasher"], r["seed"]), []).append(r)
    out = {}
    for key, rows in groups.items():
        def row(k, canon):
            for r in rows:
                if r["k"] == k and bool(r["hash_canonical"]) == canon:
                    return r
            return None

        info = {"f": None, "c": None, "rot": None, "ok": False, "why": ""}
        out[key] = info
        r1, r1c, r2 = row(1, False), row(1, True), row(2, False)
        if not (r1 and r1c and r2):
            info["why"] = "k = 1 / k = 2 rows missing"
            continue
        f, c = [None] * 4, [None] * 4
        for i, ch in enumerate(r1["text"].encode()):
            b = CODE[ch]
            if f[b] is None:
                f[b] = r1["hashes"][i]
            elif f[b] != r1["hashes"][i]:
                info["why"] = "k = 1 forward hash is not a function of the base"
        for i, ch in enumerate(r1c["text"].encode()):
            b = CODE[ch]
            v = (r1c["hashes"][i] - f[b]) & M32
            if c[b] is None:
                c[b] = v
            elif c[b] != v:
                info["why"] = "k = 1 canonical hash is not f[b] + c[b]"
        if info["why"] or None in f or None in c:
            info["why"] = info["why"] or "a base never occurs in the k = 1 text"
            continue
        t2 = r2["text"].encode()
        rots = [R for R in range(32)
                if all(table_hash(t2, i, 2, f, c, R, False) == h for i, h in enumerate(r2["hashes"]))]
        if not rots:
            info["why"] = "no rotation reproduces the k = 2 forward hashes: not a rotate-xor table hasher"
            continue
        info.update(f=f, c=c, rot=rots[0])
        bad = []
        for r in rows:
            t = r["text"].encode()
            for i, h in enumerate(r["hashes"]):
                if table_hash(t, i, r["k"], f, c, rots[0], bool(r["hash_canonical"])) != h:
                    bad.append((r["k"], bool(r["hash_canonical"]), i))
                    break
        info["ok"] = not bad
        if bad:
            info["why"] = f"tables from k = 1 do not reproduce k, canonical, index = {bad[:4]}"
    return out


def pack_ascii_n(text: bytes):
    """(packed 2-bit codes, ambiguity bit mask, both with 16 bytes of padding)."""
    a = np.frombuffer(text, dtype=np.uint8)
    codes = (a >> 1) & 3
    n = len(a)
    pad = np.zeros((n + 3) // 4 * 4, dtype=np.uint8)
    pad[:n] = codes
    q = pad.reshape(-1, 4)
    packed = np.zeros((n + 3) // 4 + 16, dtype=np.uint8)
    packed[:q.shape[0]] = q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)
    up = a & 0xDF
    bits = ~np.isin(up, np.frombuffer(b"ACGT", dtype=np.uint8))
    amb = np.zeros((n + 7) // 8 + 16, dtype=np.uint8)
    pk = np.packbits(bits.astype(np.uint8), bitorder="little")
    amb[:pk.size] = pk
    return packed, amb


def expected_case(o, dump, case, tables):
    """Oracle outputs for one dumped case, computed with the tables derived from the dump."""
    t = tables[(case["hasher"], case["seed"])]
    h = o.make_hasher_tables(t["f"], t["c"], t["rot"], bool(case["hash_canonical"]))
    pr = o.make_params(case["k"], case["w"], canonical=bool(case["builder_canonical"]), mode=case["mode"], hasher=h)
    text = dump[case["seq"]].encode()
    packed, amb = pack_ascii_n(text)
    n = len(text)
    if case["seq"] == "nseq":
        pos = o.run_skip_ambiguous(packed, 0, n, amb, 0, pr)
        sk = None
    else:
        pos, sk = o.run(packed, 0, n, pr, "stream", want_sk=(case["mode"] == 0))
    length = case["k"] if case["mode"] == 0 else case["k"] + case["w"] - 1
    v64 = o.values_u64(packed, 0, length, bool(case["builder_canonical"]), pos) if length <= 32 else None
    v128 = o.values_u128(packed, 0, length, bool(case["builder_canonical"]), pos) if length <= 64 else None
    return packed, amb, n, pos, sk, v64, v128


def check_case_against(case, pos, sk, v64, v128, who: str):
    name = case["name"]
    assert list(map(int, pos)) == case["pos"], f"{who}: positions differ from the reference for {name}"
    if case.get("sk") is not None and sk is not None:
        assert list(map(int, sk)) == case["sk"], f"{who}: super-k-mer starts differ from the reference for {name}"
    if case.get("values_u64") is not None and v64 is not None:
        assert list(map(int, v64)) == case["values_u64"], f"{who}: u64 values differ from the reference for {name}"
    if case.get("values_u128") is not None and v128 is not None:
        got = [str(int(lo) | (int(hi) << 64)) for lo, hi in v128]
        assert got == case["values_u128"], f"{who}: u128 values differ from the reference for {name}"


def synth_dump_from_oracle(o) -> dict:
    """A dump in the generator's format, but produced by the ORACLE (self-test of this checker and
    of the GPU path's table-hasher plumbing; says nothing about the reference)."""
    seq = lcg_dna(4096, 42)
    nseq = bytearray(seq)
    for start, ln in [(100, 1), (300, 3), (700, 60), (1500, 1), (1501, 1), (2000, 250), (3900, 20), (4090, 6)]:
        nseq[start:start + ln] = b"N" * ln
    dump = {"generator": "tests/refdump.py:synth_dump_from_oracle (NOT the reference)", "seq": seq.decode(),
            "nseq": bytes(nseq).decode(), "hashes": [], "cases": []}
    hashers = {("nt", None): o.make_hasher("nt", True), ("mul", None): o.make_hasher("mul", True)}
    # a made-up "seeded" table hasher and a non-table hasher stand-in
    hashers[("nt", 1234)] = o.make_hasher_tables([0x1234567, 0x89abcdef, 0x0f1e2d3c, 0x4b5a6978],
                                                 [0x0f1e2d3c, 0x4b5a6978, 0x1234567, 0x89abcdef], 7, True)
    for (name, seed), h in hashers.items():
        f, c, rot = list(h.f), list(h.c), h.rot
        for canon in (False, True):
            for k in (1, 2, 3, 5, 8, 16, 31, 32, 33, 47):
                text = b"ACTGACTGGTCA" + seq
                hs = [table_hash(text, i, k, f, c, rot, canon) for i in range(64)]
                dump["hashes"].append({"hasher": name, "hash_canonical": canon, "seed": seed, "k": k,
                                       "text": text[:64 + k - 1].decode(), "hashes": hs})
    for k in (1, 2, 5, 31):  # "antilex": deliberately not a rotate-xor table hash
        text = b"ACTGACTGGTCA" + seq
        for canon in (False, True):
            hs = [(sum(CODE[text[i + j]] * 977 ** j for j in range(k)) * 2654435761) & M32 for i in range(64)]
            dump["hashes"].append({"hasher": "antilex", "hash_canonical": canon, "seed": None, "k": k,
                                   "text": text[:64 + k - 1].decode(), "hashes": hs})
    tables = derive_tables(dump["hashes"])

    def add(name, hasher, seed, hash_canon, builder_canon, mode, k, w, seqname):
        case = {"name": name, "hasher": hasher, "hash_canonical": hash_canon, "seed": seed,
                "builder_canonical": builder_canon, "mode": mode, "k": k, "w": w, "seq": seqname}
        _, _, _, pos, sk, v64, v128 = expected_case(o, dump, case, tables)
        case["pos"] = [int(x) for x in pos]
        case["sk"] = [int(x) for x in sk] if sk is not None else None
        length = k if mode == 0 else k + w - 1
        case["values_u64"] = [int(x) for x in v64] if (v64 is not None and seqname == "seq" and hash_canon == builder_canon) else None
        case["values_u128"] = ([str(int(lo) | (int(hi) << 64)) for lo, hi in v128]
                               if (length > 32 and v128 is not None and seqname == "seq") else None)
        dump["cases"].append(case)

    for (k, w) in ((31, 19), (21, 11), (5, 7), (32, 2), (13, 33), (8, 101)):
        for hn in ("nt", "mul"):
            add(f"{hn}_fwd_k{k}_w{w}", hn, None, False, False, 0, k, w, "seq")
            add(f"{hn}_fwdbuilder_canhash_k{k}_w{w}", hn, None, True, False, 0, k, w, "seq")
            if (k + w - 1) % 2 == 1:
                add(f"{hn}_can_k{k}_w{w}", hn, None, True, True, 0, k, w, "seq")
    add("nt_seed_can_k31_w19", "nt", 1234, True, True, 0, 31, 19, "seq")
    for (k, w) in ((31, 11), (9, 5)):
        add(f"nt_can_closed_k{k}_w{w}", "nt", None, True, True, 1, k, w, "seq")
        add(f"nt_can_open_k{k}_w{w}", "nt", None, True, True, 2, k, w, "seq")
        add(f"nt_fwd_closed_k{k}_w{w}", "nt", None, False, False, 1, k, w, "seq")
    for (k, w) in ((31, 19), (21, 11), (5, 7)):
        add(f"nt_can_skipamb_k{k}_w{w}", "nt", None, True, True, 0, k, w, "nseq")
        add(f"nt_can_closed_skipamb_k{k}_w{w}", "nt", None, True, True, 1, k, w, "nseq")
    return dump
