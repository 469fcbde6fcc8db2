"""ctypes binding of the CPU ORACLE (oracle/libmzoracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
See oracle/mzoracle.h for the parity status of each stage.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmzoracle.so")

MINIMIZER, CLOSED_SYNCMER, OPEN_SYNCMER = 0, 1, 2


class Hasher(C.Structure):
    _fields_ = [("f", C.c_uint32 * 4), ("c", C.c_uint32 * 4), ("rot", C.c_uint32),
                ("canonical", C.c_uint32)]


class Params(C.Structure):
    _fields_ = [("k", C.c_uint32), ("w", C.c_uint32), ("strand_tiebreak", C.c_uint32),
                ("mode", C.c_uint32), ("hasher", Hasher)]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (seconds)."""
    srcs = [os.path.join(_HERE, f) for f in ("mzoracle.c", "mzoracle.h", "mzbaseline_avx2.c", "Makefile")]
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(f) for f in srcs)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmzoracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        u8p, u32p, u64p = C.c_void_p, C.c_void_p, C.c_void_p
        L.mzo_hasher_nt.argtypes = [C.POINTER(Hasher), C.c_int]
        L.mzo_hasher_mul.argtypes = [C.POINTER(Hasher), C.c_int]
        L.mzo_pack_ascii.argtypes = [C.c_char_p, C.c_uint64, u8p]
        L.mzo_pack_ascii.restype = C.c_uint64
        L.mzo_revcomp.argtypes = [u8p, C.c_uint64, C.c_uint64, u8p]
        L.mzo_hash_kmer.argtypes = [u8p, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(Hasher)]
        L.mzo_hash_kmer.restype = C.c_uint32
        for fn in (L.mzo_window_positions_naive, L.mzo_window_positions_stream):
            fn.argtypes = [u8p, C.c_uint64, C.c_uint64, C.POINTER(Params), u32p]
            fn.restype = C.c_uint64
        L.mzo_collect_dedup.argtypes = [u32p, C.c_uint64, u32p, u32p]
        L.mzo_collect_dedup.restype = C.c_uint64
        L.mzo_collect_syncmers.argtypes = [u32p, C.c_uint64, C.c_uint32, C.c_int, u32p]
        L.mzo_collect_syncmers.restype = C.c_uint64
        L.mzo_run.argtypes = [u8p, C.c_uint64, C.c_uint64, C.POINTER(Params), C.c_int, u32p, u32p]
        L.mzo_run.restype = C.c_uint64
        L.mzo_run_range.argtypes = [u8p, C.c_uint64, C.c_uint64, C.POINTER(Params), C.c_uint64,
                                    C.c_uint64, u32p, u32p, C.c_uint64]
        L.mzo_run_range.restype = C.c_uint64
        L.mzo_run_mt.argtypes = [u8p, C.c_uint64, C.c_uint64, C.POINTER(Params), C.c_int, u32p,
                                 u32p, u64p, C.c_uint64]
        L.mzo_run_mt.restype = C.c_uint64
        L.mzo_values_u64.argtypes = [u8p, C.c_uint64, C.c_uint32, C.c_int, u32p, C.c_uint64, u64p]
        L.mzo_values_u128.argtypes = [u8p, C.c_uint64, C.c_uint32, C.c_int, u32p, C.c_uint64, u64p]
        L.mzo_synth_packed.argtypes = [C.c_uint64, C.c_uint64, u8p]
        L.mzo_pack_ascii_n.argtypes = [C.c_char_p, C.c_uint64, u8p, u8p]
        L.mzo_pack_ascii_n.restype = C.c_uint64
        L.mzo_collect_dedup_skip_max.argtypes = [u32p, C.c_uint64, u32p]
        L.mzo_collect_dedup_skip_max.restype = C.c_uint64
        L.mzo_run_skip_ambiguous.argtypes = [u8p, C.c_uint64, C.c_uint64, u8p, C.c_uint64,
                                             C.POINTER(Params), C.c_int, u32p]
        L.mzo_run_skip_ambiguous.restype = C.c_uint64
        L.mzo_run_reads.argtypes = [u8p, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(Params), C.c_int,
                                    u64p, u32p, u32p, u64p, C.c_uint64]
        L.mzo_run_reads.restype = C.c_uint64
        L.mzb_run_mt.argtypes = L.mzo_run_mt.argtypes
        L.mzb_run_mt.restype = C.c_uint64
        L.mzb_run_mt_slices.argtypes = [u8p, C.c_uint64, C.c_uint64, C.POINTER(Params), C.c_int, u32p,
                                        u32p, u64p, C.c_uint64, u64p, u64p]
        L.mzb_run_mt_slices.restype = C.c_uint64
        L.mzb_have_avx2.restype = C.c_int
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


ERR = (1 << 64) - 1


def make_hasher(kind: str = "nt", canonical: bool = True) -> Hasher:
    h = Hasher()
    if kind == "nt":
        lib().mzo_hasher_nt(C.byref(h), int(canonical))
    elif kind == "mul":
        lib().mzo_hasher_mul(C.byref(h), int(canonical))
    else:
        raise ValueError(kind)
    return h


def make_hasher_tables(f, c, rot: int, canonical: bool) -> Hasher:
    """Any table hasher (seeded hashers cross the product's ABI the same way)."""
    h = Hasher()
    for b in range(4):
        h.f[b], h.c[b] = int(f[b]) & 0xffffffff, int(c[b]) & 0xffffffff
    h.rot, h.canonical = int(rot), int(canonical)
    return h


def make_params(k, w, *, canonical: bool, mode=MINIMIZER, hasher: Hasher | None = None) -> Params:
    """canonical: the *builder's* CANONICAL flag (strand-aware tie-break).  The default hasher
    is NtHasher<CANONICAL> (src/lib.rs:240-321)."""
    if hasher is None:
        hasher = make_hasher("nt", canonical)
    p = Params()
    p.k, p.w, p.strand_tiebreak, p.mode, p.hasher = k, w, int(canonical), mode, hasher
    return p


def pack_ascii(seq: bytes) -> np.ndarray:
    """Pack ASCII bases into 2 bits, 4 per byte (+16 bytes of zero padding)."""
    out = np.zeros((len(seq) + 3) // 4 + 16, dtype=np.uint8)
    lib().mzo_pack_ascii(seq, len(seq), _ptr(out))
    return out


def revcomp(packed: np.ndarray, off: int, n: int) -> np.ndarray:
    out = np.zeros((n + 3) // 4 + 16, dtype=np.uint8)
    lib().mzo_revcomp(_ptr(packed), off, n, _ptr(out))
    return out


def synth_packed(seed: int, n_bases: int) -> np.ndarray:
    out = np.zeros((n_bases + 3) // 4 + 16, dtype=np.uint8)
    lib().mzo_synth_packed(seed, n_bases, _ptr(out))
    return out


def hash_kmer(packed, off, i, k, hasher: Hasher) -> int:
    return lib().mzo_hash_kmer(_ptr(packed), off, i, k, C.byref(hasher))


def window_positions(packed, off, n, params: Params, algo: str = "stream") -> np.ndarray:
    l = params.k + params.w - 1
    nwin = max(0, n - l + 1)
    out = np.zeros(max(nwin, 1), dtype=np.uint32)
    fn = lib().mzo_window_positions_naive if algo == "naive" else lib().mzo_window_positions_stream
    got = fn(_ptr(packed), off, n, C.byref(params), _ptr(out))
    assert got == nwin
    return out[:nwin]


def collect_dedup(win_pos: np.ndarray):
    win_pos = np.ascontiguousarray(win_pos, dtype=np.uint32)
    pos = np.zeros(max(len(win_pos), 1), dtype=np.uint32)
    sk = np.zeros(max(len(win_pos), 1), dtype=np.uint32)
    m = lib().mzo_collect_dedup(_ptr(win_pos), len(win_pos), _ptr(pos), _ptr(sk))
    return pos[:m].copy(), sk[:m].copy()


def collect_syncmers(win_pos: np.ndarray, w: int, open_: bool):
    win_pos = np.ascontiguousarray(win_pos, dtype=np.uint32)
    out = np.zeros(max(len(win_pos), 1), dtype=np.uint32)
    m = lib().mzo_collect_syncmers(_ptr(win_pos), len(win_pos), w, int(open_), _ptr(out))
    return out[:m].copy()


def run(packed, off, n, params: Params, algo: str = "stream", want_sk: bool = False):
    """Returns (pos, sk or None).  Raises ValueError on parameters the reference asserts on."""
    l = params.k + params.w - 1
    nwin = max(0, n - l + 1)
    pos = np.zeros(max(nwin, 1), dtype=np.uint32)
    sk = np.zeros(max(nwin, 1), dtype=np.uint32) if want_sk else None
    m = lib().mzo_run(_ptr(packed), off, n, C.byref(params), 0 if algo == "naive" else 1,
                      _ptr(pos), _ptr(sk))
    if m == ERR:
        raise ValueError("oracle: invalid parameters")
    return pos[:m].copy(), (sk[:m].copy() if want_sk else None)


def run_range(packed, off, n, params: Params, wb: int, we: int, want_sk: bool = False):
    cap = max(we - wb, 1)
    pos = np.zeros(cap, dtype=np.uint32)
    sk = np.zeros(cap, dtype=np.uint32) if want_sk else None
    m = lib().mzo_run_range(_ptr(packed), off, n, C.byref(params), wb, we, _ptr(pos), _ptr(sk), cap)
    if m == ERR:
        raise ValueError("oracle: invalid parameters")
    return pos[:m].copy(), (sk[:m].copy() if want_sk else None)


def run_mt(packed, off, n, params: Params, threads: int, want_sk=False, want_val=False,
           cap: int | None = None):
    l = params.k + params.w - 1
    nwin = max(0, n - l + 1)
    if cap is None:
        cap = max(nwin, 1)
    pos = np.empty(cap, dtype=np.uint32)
    sk = np.empty(cap, dtype=np.uint32) if want_sk else None
    val = np.empty(cap, dtype=np.uint64) if want_val else None
    m = lib().mzo_run_mt(_ptr(packed), off, n, C.byref(params), threads, _ptr(pos), _ptr(sk),
                         _ptr(val), cap)
    if m == ERR:
        raise ValueError("oracle: run_mt failed (parameters or capacity)")
    return pos[:m], (sk[:m] if want_sk else None), (val[:m] if want_val else None)


def run_reads(packed, n_reads, stride_bytes, read_len, params: Params, threads: int = 1,
              want_sk=False, want_val=False, bufs=None):
    """The per-read caller loop (bench/src/bin/paper.rs:98-105) over fixed-stride reads: CSR
    (offsets, pos, sk, val).  ``bufs`` = pre-allocated (offsets, pos, sk, val) arrays (timed runs
    allocate and touch them outside the timed region)."""
    l = params.k + params.w - 1
    nwin = max(0, read_len - l + 1)
    if bufs is None:
        cap = max(nwin * n_reads, 1)
        bufs = (np.zeros(n_reads + 1, dtype=np.uint64), np.zeros(cap, dtype=np.uint32),
                np.zeros(cap, dtype=np.uint32) if want_sk else None,
                np.zeros(cap, dtype=np.uint64) if want_val else None)
    offs, pos, sk, val = bufs
    m = lib().mzo_run_reads(_ptr(packed), n_reads, stride_bytes, read_len, C.byref(params), threads,
                            _ptr(offs), _ptr(pos), _ptr(sk), _ptr(val), len(pos))
    if m == ERR:
        raise ValueError("oracle: run_reads failed (parameters or capacity)")
    return offs, pos[:m], (sk[:m] if sk is not None else None), (val[:m] if val is not None else None)


def baseline_run_mt(packed, off, n, params: Params, threads: int, want_sk=False, want_val=False,
                    cap: int | None = None):
    """8-lane AVX2 + pthreads CPU baseline (oracle/mzbaseline_avx2.c); scalar fallback inside."""
    l = params.k + params.w - 1
    nwin = max(0, n - l + 1)
    if cap is None:
        cap = max(nwin, 1)
    pos = np.empty(cap, dtype=np.uint32)
    sk = np.empty(cap, dtype=np.uint32) if want_sk else None
    val = np.empty(cap, dtype=np.uint64) if want_val else None
    m = lib().mzb_run_mt(_ptr(packed), off, n, C.byref(params), threads, _ptr(pos), _ptr(sk),
                         _ptr(val), cap)
    if m == ERR:
        raise ValueError("baseline: run failed (parameters or capacity)")
    return pos[:m], (sk[:m] if want_sk else None), (val[:m] if want_val else None)


def baseline_run_mt_slices(packed, off, n, params: Params, threads: int, pos, sk, val):
    """Timed CPU arm: every thread writes into its own slice of the pre-allocated arrays
    (pos / sk / val, sk and val may be None); returns (total, starts, counts) -- the logical
    output is the concatenation of pos[starts[t]:starts[t]+counts[t]]."""
    starts = np.zeros(threads, dtype=np.uint64)
    counts = np.zeros(threads, dtype=np.uint64)
    m = lib().mzb_run_mt_slices(_ptr(packed), off, n, C.byref(params), threads, _ptr(pos), _ptr(sk),
                                _ptr(val), len(pos), _ptr(starts), _ptr(counts))
    if m == ERR:
        raise ValueError("baseline: run failed (parameters or slice capacity)")
    return int(m), starts, counts


def have_avx2() -> bool:
    return bool(lib().mzb_have_avx2())


def values_u64(packed, off, length, canonical, pos: np.ndarray) -> np.ndarray:
    pos = np.ascontiguousarray(pos, dtype=np.uint32)
    out = np.zeros(max(len(pos), 1), dtype=np.uint64)
    lib().mzo_values_u64(_ptr(packed), off, length, int(canonical), _ptr(pos), len(pos), _ptr(out))
    return out[:len(pos)]


def values_u128(packed, off, length, canonical, pos: np.ndarray) -> np.ndarray:
    """Returns an (m, 2) uint64 array of (lo, hi)."""
    pos = np.ascontiguousarray(pos, dtype=np.uint32)
    out = np.zeros((max(len(pos), 1), 2), dtype=np.uint64)
    lib().mzo_values_u128(_ptr(packed), off, length, int(canonical), _ptr(pos), len(pos), _ptr(out))
    return out[:len(pos)]


SKIPPED = 0xFFFFFFFE


def pack_ascii_n(seq: bytes):
    """ASCII -> (packed 2-bit codes, ambiguity bit mask); everything but ACGTacgt is ambiguous."""
    packed = np.zeros((len(seq) + 3) // 4 + 16, dtype=np.uint8)
    amb = np.zeros((len(seq) + 7) // 8 + 16, dtype=np.uint8)
    lib().mzo_pack_ascii_n(seq, len(seq), _ptr(packed), _ptr(amb))
    return packed, amb


def collect_dedup_skip_max(win_pos: np.ndarray) -> np.ndarray:
    win_pos = np.ascontiguousarray(win_pos, dtype=np.uint32)
    pos = np.zeros(max(len(win_pos), 1), dtype=np.uint32)
    m = lib().mzo_collect_dedup_skip_max(_ptr(win_pos), len(win_pos), _ptr(pos))
    return pos[:m].copy()


def run_skip_ambiguous(packed, off, n, amb, amb_off, params: Params, algo: str = "stream"):
    """run_skip_ambiguous_windows (src/lib.rs:451-496): positions only."""
    l = params.k + params.w - 1
    nwin = max(0, n - l + 1)
    pos = np.zeros(max(nwin, 1), dtype=np.uint32)
    m = lib().mzo_run_skip_ambiguous(_ptr(packed), off, n, _ptr(amb), amb_off, C.byref(params),
                                     0 if algo == "naive" else 1, _ptr(pos))
    if m == ERR:
        raise ValueError("oracle: invalid parameters")
    return pos[:m].copy()
