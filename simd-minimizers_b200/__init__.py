"""simd-minimizers_b200 -- B200-native random-minimizer path behind the simd-minimizers API.

Host-side mirror (Python, over the C ABI in include/mz_b200.h) of the reference crate's public
interface for this path (rust-seq/simd-minimizers v3.0.0, src/lib.rs:225-654):

    minimizer_positions(seq, k, w)              canonical_minimizer_positions(seq, k, w)
    minimizers(k, w) / canonical_minimizers(k, w)
    closed_syncmers / canonical_closed_syncmers / open_syncmers / canonical_open_syncmers
    canonical_syncmers (README name; alias of canonical_closed_syncmers)
    Builder.hasher(h).super_kmers(sk).run(seq, pos) -> Output
    Output.values_u64() / values_u128() / pos_and_values_u64() / pos_and_values_u128()

Same names, argument meaning and error behaviour (the reference's assert!/panic! become
``AssertionError`` with the reference's message).  All compute happens in hand-written
sm_100a kernels inside libmzb200.so; nothing here computes a minimizer on the CPU.
The package directory name contains a '-', so import it with
``importlib.import_module("simd-minimizers_b200")``.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _ffi
from ._ffi import MzError, MzOut, MzParams, MzTiming

__all__ = [
    "PackedSeq", "PackedSeqVec", "AsciiSeq", "BitSeq", "PackedNSeq", "PackedNSeqVec", "NtHasher", "MulHasher", "U32Vec", "Builder", "Output",
    "minimizers", "canonical_minimizers", "closed_syncmers", "canonical_closed_syncmers",
    "open_syncmers", "canonical_open_syncmers", "canonical_syncmers", "minimizer_positions",
    "canonical_minimizer_positions", "Context", "default_context", "MzError",
]

_PANICS = {2, 3, 4, 5, 6, 7}  # codes that are assert!/panic! in the reference


def _check(code: int):
    if code in _PANICS:
        raise AssertionError(_ffi.lib().mz_strerror(code).decode())
    _ffi.check(code)


# ------------------------------------------------------------------------------------------
# packed-seq stand-ins: just enough of PackedSeq / PackedSeqVec for this path
# (packed-seq 5.0.0 is an external crate; layout: 4 bases per byte, first base in the low
#  bits, A=0 C=1 T=2 G=3, see src/lib.rs:120-123)
# ------------------------------------------------------------------------------------------
_PAD = 16


class PackedSeq:
    """Borrowed view: bases [offset, offset+len) of a packed byte buffer."""

    def __init__(self, data: np.ndarray, offset: int, length: int):
        assert data.dtype == np.uint8 and data.flags.c_contiguous
        assert (offset + length + 3) // 4 <= data.size
        self.data, self.offset, self.len = data, int(offset), int(length)

    def __len__(self):
        return self.len

    def slice(self, start: int, end: int) -> "PackedSeq":
        assert 0 <= start <= end <= self.len
        return PackedSeq(self.data, self.offset + start, end - start)

    def as_slice(self) -> "PackedSeq":
        return self

    def get(self, i: int) -> int:
        p = self.offset + i
        return (int(self.data[p >> 2]) >> (2 * (p & 3))) & 3

    def to_revcomp(self) -> "PackedSeqVec":
        p = np.arange(self.offset, self.offset + self.len, dtype=np.int64)
        codes = (self.data[p >> 2] >> (2 * (p & 3)).astype(np.uint8)) & 3
        return PackedSeqVec.from_codes((codes[::-1] ^ 2).astype(np.uint8))


class PackedSeqVec:
    """Owning packed sequence."""

    def __init__(self, data: np.ndarray, length: int):
        self.data, self.len = data, int(length)

    @staticmethod
    def from_codes(codes: np.ndarray) -> "PackedSeqVec":
        n = int(codes.size)
        padded = np.zeros((n + 3) // 4 * 4, dtype=np.uint8)
        padded[:n] = codes
        q = padded.reshape(-1, 4)
        data = np.zeros((n + 3) // 4 + _PAD, dtype=np.uint8)
        data[:q.shape[0]] = q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)
        return PackedSeqVec(data, n)

    @staticmethod
    def from_ascii(seq: bytes) -> "PackedSeqVec":
        a = np.frombuffer(bytes(seq), dtype=np.uint8)
        return PackedSeqVec.from_codes((a >> 1) & 3)

    @staticmethod
    def random(n: int, seed: int | None = None) -> "PackedSeqVec":
        rng = np.random.default_rng(seed)
        data = np.zeros((n + 3) // 4 + _PAD, dtype=np.uint8)
        data[:(n + 3) // 4] = rng.integers(0, 256, size=(n + 3) // 4, dtype=np.uint8)
        return PackedSeqVec(data, n)

    def __len__(self):
        return self.len

    def as_slice(self) -> PackedSeq:
        return PackedSeq(self.data, 0, self.len)

    def slice(self, start: int, end: int) -> PackedSeq:
        return self.as_slice().slice(start, end)


class AsciiSeq:
    """ASCII DNA text (packed_seq::AsciiSeq): characters map to 2-bit codes with (c >> 1) & 3,
    so ACTGactg work and anything else aliases silently, as in the reference.  Packed on the
    device (mz_run_ascii); never packed on the host."""

    def __init__(self, seq: bytes):
        self.seq = bytes(seq)
        self.len = len(self.seq)

    def __len__(self):
        return self.len

    def as_slice(self) -> "AsciiSeq":
        return self

    def slice(self, start: int, end: int) -> "AsciiSeq":
        return AsciiSeq(self.seq[start:end])

    def pack(self, ctx: "Context | None" = None) -> "PackedSeqVec":
        """PackedSeqVec::from_ascii, computed on the device (mz_pack_ascii)."""
        data = np.zeros((self.len + 3) // 4 + _PAD, dtype=np.uint8)
        ctx = ctx or default_context()
        _check(_ffi.lib().mz_pack_ascii(ctx.handle, self.seq, self.len, data.ctypes.data))
        return PackedSeqVec(data, self.len)


class BitSeq:
    """One bit per base (packed_seq::BitSeq): base i -> bit (offset+i)&7 of byte (offset+i)>>3."""

    def __init__(self, data: np.ndarray, offset: int, length: int):
        assert data.dtype == np.uint8 and data.flags.c_contiguous
        assert (offset + length + 7) // 8 <= data.size
        self.data, self.offset, self.len = data, int(offset), int(length)

    def get(self, i: int) -> int:
        p = self.offset + i
        return (int(self.data[p >> 3]) >> (p & 7)) & 1


class PackedNSeq:
    """packed_seq::PackedNSeq: 2-bit codes + ambiguity mask (the codes under ambiguous bases are
    arbitrary and never influence the output)."""

    def __init__(self, seq: PackedSeq, ambiguous: BitSeq):
        assert seq.len == ambiguous.len
        self.seq, self.ambiguous, self.len = seq, ambiguous, seq.len

    def __len__(self):
        return self.len

    def as_slice(self) -> "PackedNSeq":
        return self

    def slice(self, start: int, end: int) -> "PackedNSeq":
        assert 0 <= start <= end <= self.len
        return PackedNSeq(self.seq.slice(start, end),
                          BitSeq(self.ambiguous.data, self.ambiguous.offset + start, end - start))


class PackedNSeqVec:
    """Owning PackedNSeq.  ``from_ascii`` runs on the device (mz_pack_ascii_n): every character
    outside ACGTacgt is ambiguous."""

    def __init__(self, seq: PackedSeqVec, amb: np.ndarray):
        self.seq, self.amb, self.len = seq, amb, seq.len

    @staticmethod
    def from_ascii(seq: bytes, ctx: "Context | None" = None) -> "PackedNSeqVec":
        seq = bytes(seq)
        n = len(seq)
        data = np.zeros((n + 3) // 4 + _PAD, dtype=np.uint8)
        amb = np.zeros((n + 7) // 8 + _PAD, dtype=np.uint8)
        ctx = ctx or default_context()
        _check(_ffi.lib().mz_pack_ascii_n(ctx.handle, seq, n, data.ctypes.data, amb.ctypes.data))
        return PackedNSeqVec(PackedSeqVec(data, n), amb)

    @staticmethod
    def from_parts(seq: PackedSeqVec, amb_bits: np.ndarray) -> "PackedNSeqVec":
        """amb_bits: one 0/1 entry per base."""
        assert amb_bits.size == seq.len
        amb = np.zeros((seq.len + 7) // 8 + _PAD, dtype=np.uint8)
        pk = np.packbits(amb_bits.astype(np.uint8), bitorder="little")
        amb[:pk.size] = pk
        return PackedNSeqVec(seq, amb)

    def __len__(self):
        return self.len

    def as_slice(self) -> PackedNSeq:
        return PackedNSeq(self.seq.as_slice(), BitSeq(self.amb, 0, self.len))

    def slice(self, start: int, end: int) -> PackedNSeq:
        return self.as_slice().slice(start, end)


# ------------------------------------------------------------------------------------------
# hashers (seq-hash 0.2.0 NtHasher<RC> / MulHasher<RC>, passed to the device as tables)
# ------------------------------------------------------------------------------------------
class _TableHasher:
    _setter = None

    def __init__(self, k: int, canonical: bool = True):
        self._k, self._canonical = int(k), bool(canonical)

    @classmethod
    def new(cls, k: int, canonical: bool = True):
        return cls(k, canonical)

    def is_canonical(self) -> bool:
        return self._canonical

    @classmethod
    def new_with_seed(cls, k: int, seed: int, canonical: bool = True):
        """``H::new_with_seed(k, seed)`` (src/lib.rs:157, src/test.rs:287).  seq-hash 0.2.0 derives
        the per-base tables from the seed inside the un-vendored crate, so the arithmetic cannot
        be restated here; the Rust shim reads the tables out of the hasher object.  Pass them with
        ``from_tables`` instead."""
        raise NotImplementedError(
            "seeded tables live in seq-hash (not in the reference tree): use "
            f"{cls.__name__}.from_tables(k, f, c, rot, canonical) with the tables of the seeded hasher")

    @classmethod
    def from_tables(cls, k: int, f, c, rot: int = 7, canonical: bool = True):
        """Any table hasher (seeded NtHasher / MulHasher): fw = XOR rotl(f[b], rot*(k-1-j)),
        rc = XOR rotl(c[b], rot*j), tables indexed by packed code A=0 C=1 T=2 G=3."""
        h = cls(k, canonical)
        h._tables = ([int(x) & 0xffffffff for x in f], [int(x) & 0xffffffff for x in c], int(rot))
        assert len(h._tables[0]) == 4 and len(h._tables[1]) == 4
        return h

    def k(self) -> int:
        return self._k

    def _apply(self, p: MzParams):
        t = getattr(self, "_tables", None)
        if t is not None:
            f, c, rot = t
            _check(_ffi.lib().mz_params_set_tables(C.byref(p), (C.c_uint32 * 4)(*f), (C.c_uint32 * 4)(*c),
                                                   rot, int(self._canonical)))
            return
        _check(getattr(_ffi.lib(), self._setter)(C.byref(p), int(self._canonical)))


class NtHasher(_TableHasher):
    _setter = "mz_params_set_nthash"


class MulHasher(_TableHasher):
    """seq-hash MulHasher.  The tables behind mz_params_set_mulhash are a restatement that no
    vector of the reference pins: GPU == oracle is tested, reference parity is
    open until tests/golden/reference_dump.json is generated with the crate."""
    _setter = "mz_params_set_mulhash"


# ------------------------------------------------------------------------------------------
# context
# ------------------------------------------------------------------------------------------
class Context:
    """Streams + device scratch for a set of GPUs (mz_ctx).  Not thread-safe: one per thread,
    like the reference's thread_local CACHE (src/lib.rs:217-219)."""

    def __init__(self, devices: list[int] | None = None):
        self._h = C.c_void_p()
        if devices:
            arr = (C.c_int * len(devices))(*devices)
            _check(_ffi.lib().mz_ctx_create(arr, len(devices), C.byref(self._h)))
        else:
            _check(_ffi.lib().mz_ctx_create(None, 0, C.byref(self._h)))

    def close(self):
        if self._h:
            _ffi.lib().mz_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def device_count(self) -> int:
        return _ffi.lib().mz_ctx_device_count(self._h)

    def last_timing(self) -> dict:
        t = MzTiming()
        _check(_ffi.lib().mz_last_timing(self._h, C.byref(t)))
        return {"h2d_ms": t.h2d_ms, "kernel_ms": t.kernel_ms, "d2h_ms": t.d2h_ms,
                "total_ms": t.total_ms, "kernel_launches": t.kernel_launches}


_tls = threading.local()


def default_context() -> Context:
    ctx = getattr(_tls, "ctx", None)
    if ctx is None:
        ctx = _tls.ctx = Context()
    return ctx


# ------------------------------------------------------------------------------------------
# Vec<u32> stand-in with append semantics
# ------------------------------------------------------------------------------------------
class U32Vec:
    """Growable u32 vector (the caller's ``Vec<u32>``; results are *appended*, src/lib.rs:80)."""

    def __init__(self, init=None):
        self._a = np.zeros(0, dtype=np.uint32) if init is None else np.array(init, dtype=np.uint32)

    def clear(self):
        self._a = self._a[:0]

    def __len__(self):
        return int(self._a.size)

    def __iter__(self):
        return iter(self._a.tolist())

    def __getitem__(self, i):
        return self._a[i]

    def last(self):
        return int(self._a[-1]) if self._a.size else None

    @property
    def array(self) -> np.ndarray:
        return self._a

    def tolist(self):
        return self._a.tolist()

    def __eq__(self, other):
        return self.tolist() == (other.tolist() if hasattr(other, "tolist") else list(other))

    def _extend(self, arr: np.ndarray):
        self._a = np.concatenate([self._a, arr.astype(np.uint32, copy=False)])


def _as_vec(v):
    if isinstance(v, U32Vec):
        return v
    raise TypeError("positions must be collected into a simd_minimizers U32Vec")


# ------------------------------------------------------------------------------------------
# Builder / Output  (src/lib.rs:225-630)
# ------------------------------------------------------------------------------------------
class Output:
    """src/lib.rs:232-237, 579-630.  ``len`` is k for minimizers, k+w-1 for syncmers.  Values are
    lazy, as in the reference: ``run`` moves positions only, ``values_*`` computes the k-mers of
    *all* of ``min_pos`` on the device (mz_values; the sequence of the run is still resident there)."""

    def __init__(self, builder: "Builder", seq, min_pos: U32Vec):
        self.len = builder.k if builder.syncmer == 0 else builder.k + builder.w - 1
        self._b, self.seq, self.min_pos = builder, seq, min_pos

    def _values(self, bits: int) -> np.ndarray:
        if self.len > bits // 2:
            raise AssertionError(_ffi.lib().mz_strerror(7).decode())
        seq = self.seq
        if isinstance(seq, AsciiSeq):  # packed on the device, values read the packed form
            seq = seq.pack(self._b._ctx).as_slice()
        elif isinstance(seq, PackedNSeq):
            seq = seq.seq
        p = self._b._params(0)
        ctx = self._b._ctx or default_context()
        pos = np.ascontiguousarray(self.min_pos.array, dtype=np.uint32)
        vw = bits // 64
        val = np.empty(max(pos.size, 1) * vw, dtype=np.uint64)
        _check(_ffi.lib().mz_values(ctx.handle, C.byref(p), seq.data.ctypes.data, seq.offset, seq.len,
                                    pos.ctypes.data, pos.size, bits, val.ctypes.data))
        return val[:pos.size] if vw == 1 else val[:2 * pos.size].reshape(pos.size, 2)

    def values_u64(self):
        return self._values(64)

    def values_u128(self):
        v = self._values(128)
        return [int(lo) | (int(hi) << 64) for lo, hi in v]

    def pos_and_values_u64(self):
        return list(zip(self.min_pos.tolist(), self._values(64).tolist()))

    def pos_and_values_u128(self):
        return list(zip(self.min_pos.tolist(), self.values_u128()))


class Builder:
    """Type-state builder of the reference (src/lib.rs:225-230): CANONICAL, hasher, SkPos,
    SYNCMER are plain attributes here."""

    def __init__(self, k: int, w: int, canonical: bool, syncmer: int, hasher=None, sk_pos=None,
                 ctx: Context | None = None):
        self.k, self.w, self.canonical, self.syncmer = int(k), int(w), bool(canonical), int(syncmer)
        self._hasher, self._sk_pos, self._ctx = hasher, sk_pos, ctx

    # -- configuration ------------------------------------------------------------------
    def hasher(self, h) -> "Builder":
        if self._sk_pos is not None:  # src/lib.rs:323-338: hasher() only before super_kmers()
            raise TypeError("hasher() must be called before super_kmers()")
        return Builder(self.k, self.w, self.canonical, self.syncmer, h, None, self._ctx)

    def super_kmers(self, sk_pos: U32Vec) -> "Builder":
        if self.syncmer != 0:  # src/lib.rs:339: only Builder<.., 0>
            raise TypeError("super_kmers() is only available for minimizers")
        return Builder(self.k, self.w, self.canonical, 0, self._hasher, _as_vec(sk_pos), self._ctx)

    def context(self, ctx: Context) -> "Builder":
        return Builder(self.k, self.w, self.canonical, self.syncmer, self._hasher, self._sk_pos, ctx)

    # -- execution ----------------------------------------------------------------------
    def _params(self, value_bits: int) -> MzParams:
        p = MzParams()
        L = _ffi.lib()
        _check(L.mz_params_nthash(C.byref(p), self.k, self.w, self.syncmer, int(self.canonical)))
        if self._hasher is not None:
            if self._hasher.k() != self.k:
                raise AssertionError("hasher.k() must equal the builder's k")
            self._hasher._apply(p)
        p.want_sk = 1 if self._sk_pos is not None else 0
        p.value_bits = value_bits
        return p

    def _execute(self, seq, value_bits: int, skip: bool = False):
        seq = seq.as_slice()
        p = self._params(value_bits)
        L = _ffi.lib()
        n = seq.len
        _check(L.mz_params_validate(C.byref(p), n))
        l = self.k + self.w - 1
        nwin = max(0, n - l + 1)
        if self.syncmer == 0:
            dens = 2.0 / (self.w + 1)
        elif self.syncmer == 1:
            dens = 1.0 if self.w == 1 else 2.0 / self.w
        else:
            dens = 1.0 / self.w
        cap = int(min(nwin, nwin * dens * 1.25 + 4096))
        ctx = self._ctx or default_context()
        vw = value_bits // 64
        while True:
            pos = np.empty(max(cap, 1), dtype=np.uint32)
            sk = np.empty(max(cap, 1), dtype=np.uint32) if p.want_sk else None
            val = np.empty(max(cap, 1) * max(vw, 1), dtype=np.uint64) if vw else None
            out = MzOut(pos.ctypes.data, sk.ctypes.data if sk is not None else None,
                        val.ctypes.data if val is not None else None, cap, 0)
            if skip and isinstance(seq, AsciiSeq):
                rc = L.mz_run_ascii_skip_ambiguous(ctx.handle, C.byref(p), seq.seq, n, C.byref(out))
            elif skip:
                rc = L.mz_run_skip_ambiguous(ctx.handle, C.byref(p), seq.seq.data.ctypes.data,
                                             seq.seq.offset, n, seq.ambiguous.data.ctypes.data,
                                             seq.ambiguous.offset, C.byref(out))
            elif isinstance(seq, AsciiSeq):
                rc = L.mz_run_ascii(ctx.handle, C.byref(p), seq.seq, n, C.byref(out))
            else:
                rc = L.mz_run(ctx.handle, C.byref(p), seq.data.ctypes.data, seq.offset, n, C.byref(out))
            if rc == _ffi.MZ_ERR_CAPACITY:
                cap = int(out.count)
                continue
            _check(rc)
            break
        m = int(out.count)
        vals = None
        if vw == 1:
            vals = val[:m]
        elif vw == 2:
            vals = val[:2 * m].reshape(m, 2)
        return pos[:m], (sk[:m] if sk is not None else None), vals

    def run(self, seq, min_pos: U32Vec) -> Output:
        """Append positions to ``min_pos`` (and super-k-mer starts to the sk vector).  Positions
        only cross the bus; ``Output.values_*`` are lazy (src/lib.rs:598-629)."""
        min_pos = _as_vec(min_pos)
        pos, sk, _ = self._execute(seq, 0)
        # SIMD-collector quirk (src/collect.rs:257,267): the first new element is dropped when
        # it equals the caller's current last element.  Syncmers are appended verbatim
        # (src/syncmers.rs:167-169).
        if self.syncmer == 0 and len(min_pos) and pos.size and int(pos[0]) == min_pos.last():
            pos, sk = pos[1:], (sk[1:] if sk is not None else None)
        min_pos._extend(pos)
        if self._sk_pos is not None:
            self._sk_pos._extend(sk)
        return Output(self, seq.as_slice(), min_pos)

    def run_once(self, seq) -> np.ndarray:
        v = U32Vec()
        self.run(seq, v)
        return v.array

    def run_scalar(self, seq, min_pos: U32Vec) -> Output:
        """``run_scalar`` (src/lib.rs:372-378, 508-533): the same result through the reference's
        scalar collectors, which OVERWRITE the vectors from index 0 instead of appending
        (src/collect.rs:15-76, src/syncmers.rs:19-48).  Same device path here (the reference
        asserts scalar == SIMD, src/test.rs:53-84)."""
        min_pos = _as_vec(min_pos)
        pos, sk, _ = self._execute(seq, 0)
        min_pos.clear()
        min_pos._extend(pos)
        if self._sk_pos is not None and seq.as_slice().len >= self.k + self.w - 1:
            # an empty window stream leaves the index vector untouched (src/collect.rs:45-48)
            self._sk_pos.clear()
            self._sk_pos._extend(sk)
        return Output(self, seq.as_slice(), min_pos)

    def run_scalar_once(self, seq) -> np.ndarray:
        v = U32Vec()
        self.run_scalar(seq, v)
        return v.array

    def run_with_values(self, seq, value_bits: int = 64, skip_ambiguous: bool = False):
        """Not in the reference: positions (+ super-k-mer starts) AND values in one fused launch
        (mz_params.value_bits), for callers that always consume the values (bench.py's pos+vals
        metric).  Returns (pos, sk, vals); u128 values come as an (n, 2) array of (lo, hi)."""
        return self._execute(seq, value_bits, skip=skip_ambiguous)

    def bucket_stats(self, seq, n_buckets: int):
        """Not in the reference crate: super-k-mers (bench/src/minimizer.rs:3-36) sharded by their
        minimizer on the device (mz_run_bucket_stats).  Returns (superkmers[n_buckets],
        windows[n_buckets], n_minimizers); only the histograms leave the GPU."""
        if self.syncmer != 0:
            raise TypeError("bucket_stats() is only available for minimizers")
        seq = seq.as_slice()
        p = self._params(0)
        ctx = self._ctx or default_context()
        cnt = np.zeros(n_buckets, dtype=np.uint64)
        win = np.zeros(n_buckets, dtype=np.uint64)
        total = C.c_uint64(0)
        _check(_ffi.lib().mz_run_bucket_stats(ctx.handle, C.byref(p), seq.data.ctypes.data, seq.offset, seq.len,
                                              n_buckets, cnt.ctypes.data, win.ctypes.data, C.byref(total)))
        return cnt, win, int(total.value)

    def run_skip_ambiguous_windows(self, nseq, min_pos: U32Vec) -> Output:
        """src/lib.rs:451-496: windows containing an ambiguous base produce nothing.  ``nseq`` is
        a PackedNSeq(Vec), or an AsciiSeq (packed and masked on the device).  Canonical builders
        without super-k-mers only, as in the reference."""
        if not self.canonical or self._sk_pos is not None:
            raise TypeError("run_skip_ambiguous_windows needs a canonical builder without super_kmers()")
        if not isinstance(nseq, (PackedNSeq, PackedNSeqVec, AsciiSeq)):
            raise TypeError("run_skip_ambiguous_windows takes a PackedNSeq")
        min_pos = _as_vec(min_pos)
        pos, _, _ = self._execute(nseq, 0, skip=True)
        if self.syncmer == 0 and len(min_pos) and pos.size and int(pos[0]) == min_pos.last():
            pos = pos[1:]
        min_pos._extend(pos)
        return Output(self, nseq.as_slice(), min_pos)

    def run_skip_ambiguous_windows_once(self, nseq) -> np.ndarray:
        v = U32Vec()
        self.run_skip_ambiguous_windows(nseq, v)
        return v.array

    def run_batch(self, packed: np.ndarray, *, starts=None, lens=None, stride_bytes: int = 0,
                  read_len: int = 0, n_reads: int | None = None, value_bits: int | None = None):
        """Many independent reads in one launch (mz_run_batch).  Equivalent to the reference
        idiom ``for s in &seqs { v.clear(); builder.run(s, &mut v) }``
        (bench/src/bin/paper.rs:98-105) with the per-read results returned as CSR:
        ``offsets[r]:offsets[r+1]`` index ``pos`` / ``sk`` / ``vals``; positions are relative to
        the read start.  Either ``starts``+``lens`` (bases, arbitrary layout) or a fixed
        ``stride_bytes`` + ``read_len`` layout."""
        assert packed.dtype == np.uint8 and packed.flags.c_contiguous
        length = self.k if self.syncmer == 0 else self.k + self.w - 1
        if value_bits is None:
            value_bits = 64 if length <= 32 else 0
        p = self._params(value_bits)
        L = _ffi.lib()
        l = self.k + self.w - 1
        if starts is not None:
            starts = np.ascontiguousarray(starts, dtype=np.uint64)
            lens = np.ascontiguousarray(lens, dtype=np.uint32)
            n_reads = int(starts.size)
            nwin = int(np.maximum(lens.astype(np.int64) - l + 1, 0).sum())
        else:
            assert n_reads is not None
            nwin = max(0, read_len - l + 1) * n_reads
        _check(L.mz_params_validate(C.byref(p), int(lens.max()) if starts is not None and n_reads else read_len))
        dens = 2.0 / (self.w + 1) if self.syncmer == 0 else (2.0 / self.w if self.syncmer == 1 else 1.0 / self.w)
        cap = int(min(nwin, nwin * dens * 1.25 + n_reads + 4096))
        ctx = self._ctx or default_context()
        vw = value_bits // 64
        offsets = np.zeros(n_reads + 1, dtype=np.uint64)
        while True:
            pos = np.empty(max(cap, 1), dtype=np.uint32)
            sk = np.empty(max(cap, 1), dtype=np.uint32) if p.want_sk else None
            val = np.empty(max(cap, 1) * max(vw, 1), dtype=np.uint64) if vw else None
            out = MzOut(pos.ctypes.data, sk.ctypes.data if sk is not None else None,
                        val.ctypes.data if val is not None else None, cap, 0)
            rc = L.mz_run_batch(ctx.handle, C.byref(p), packed.ctypes.data, packed.size, n_reads,
                                starts.ctypes.data if starts is not None else None,
                                lens.ctypes.data if starts is not None else None,
                                stride_bytes, read_len, offsets.ctypes.data, C.byref(out))
            if rc == _ffi.MZ_ERR_CAPACITY:
                cap = int(out.count)
                continue
            _check(rc)
            break
        m = int(out.count)
        vals = None if not vw else (val[:m] if vw == 1 else val[:2 * m].reshape(m, 2))
        return offsets, pos[:m], (sk[:m] if sk is not None else None), vals


def minimizers(k: int, w: int) -> Builder:                  # src/lib.rs:240
    return Builder(k, w, False, 0)


def canonical_minimizers(k: int, w: int) -> Builder:        # src/lib.rs:250
    return Builder(k, w, True, 0)


def closed_syncmers(k: int, w: int) -> Builder:             # src/lib.rs:269
    return Builder(k, w, False, 1)


def canonical_closed_syncmers(k: int, w: int) -> Builder:   # src/lib.rs:282
    return Builder(k, w, True, 1)


def open_syncmers(k: int, w: int) -> Builder:               # src/lib.rs:301
    return Builder(k, w, False, 2)


def canonical_open_syncmers(k: int, w: int) -> Builder:     # src/lib.rs:311
    return Builder(k, w, True, 2)


def canonical_syncmers(k: int, w: int) -> Builder:
    """Name used by the reference docs (README.md:65, src/lib.rs:51) but absent from its code;
    provided as an alias of canonical_closed_syncmers."""
    return canonical_closed_syncmers(k, w)


def minimizer_positions(seq, k: int, w: int) -> np.ndarray:             # src/lib.rs:639
    return minimizers(k, w).run_once(seq)


def canonical_minimizer_positions(seq, k: int, w: int) -> np.ndarray:   # src/lib.rs:652
    return canonical_minimizers(k, w).run_once(seq)
