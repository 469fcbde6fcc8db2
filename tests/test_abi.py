"""CPU tests: the C-ABI library loads, exports every symbol include/mz_b200.h declares, and
its host-only entry points (parameter helpers / validation / error strings) behave.  No
compute calls -- those need a GPU and live in the -m gpu tests."""
import ctypes as C
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ffi(sm):
    return importlib.import_module("simd-minimizers_b200._ffi")


def test_header_symbols_exported(ffi):
    hdr = open(os.path.join(ROOT, "include", "mz_b200.h")).read()
    declared = set(re.findall(r"\b(mz_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(ffi.SYMBOLS), declared ^ set(ffi.SYMBOLS)
    L = ffi.lib()
    for s in declared:
        assert getattr(L, s) is not None
    assert L.mz_abi_version() == 1


def test_rust_sys_bindings_cover_the_header():
    """rust/mzb200-sys declares exactly the header's entry points (the crate cannot be compiled in
    this image, so at least the symbol lists must not drift apart)."""
    hdr = open(os.path.join(ROOT, "include", "mz_b200.h")).read()
    declared = set(re.findall(r"\b(mz_[a-z0-9_]+)\s*\(", hdr))
    rs = open(os.path.join(ROOT, "rust", "mzb200-sys", "src", "lib.rs")).read()
    bound = set(re.findall(r"pub fn (mz_[a-z0-9_]+)\s*\(", rs))
    assert declared == bound, declared ^ bound
    # the safe shim only calls what the -sys crate binds
    shim = open(os.path.join(ROOT, "rust", "simd-minimizers-b200", "src", "lib.rs")).read()
    used = set(re.findall(r"sys::(mz_[a-z0-9_]+)\s*\(", shim))
    assert used <= bound, used - bound


def test_struct_layout_matches_header(ffi):
    assert C.sizeof(ffi.MzParams) == 4 * (6 + 4 + 4 + 2 + 2)
    assert C.sizeof(ffi.MzOut) == 40
    assert C.sizeof(ffi.MzTiming) == 24


def test_param_helpers_and_validation(ffi):
    L = ffi.lib()
    p = ffi.MzParams()
    assert L.mz_params_nthash(C.byref(p), 31, 19, 0, 1) == 0
    assert (p.k, p.w, p.mode, p.strand_tiebreak, p.hash_canonical, p.rot) == (31, 19, 0, 1, 1, 7)
    assert list(p.f) == [0x95c60474, 0x62a02b4c, 0x82572324, 0x4be24456]
    assert list(p.c) == [p.f[2], p.f[3], p.f[0], p.f[1]]
    assert L.mz_params_validate(C.byref(p), 1000) == 0
    p.value_bits = 64
    assert L.mz_params_validate(C.byref(p), 1000) == 0
    p.k = 33
    assert L.mz_params_validate(C.byref(p), 1000) == 7      # value width
    p.value_bits = 0
    assert L.mz_params_validate(C.byref(p), 1 << 32) == 3   # too long, src/sliding_min.rs:96-99
    p.k, p.w = 4, 3
    assert L.mz_params_validate(C.byref(p), 100) == 4       # even l, src/canonical.rs:13-16
    p.w = 1 << 15
    assert L.mz_params_validate(C.byref(p), 100) == 2       # src/sliding_min.rs:92-95
    assert L.mz_params_nthash(C.byref(p), 5, 4, 2, 0) == 0
    assert L.mz_params_validate(C.byref(p), 100) == 5       # open syncmers need odd w
    assert L.mz_params_nthash(C.byref(p), 5, 7, 0, 1) == 0
    assert L.mz_params_set_nthash(C.byref(p), 0) == 0
    assert L.mz_params_validate(C.byref(p), 100) == 6       # canonical builder, forward hasher
    assert L.mz_params_mulhash(C.byref(p), 5, 7, 0, 1) == 0
    assert list(p.f) == [(b * 0x27220a95) & 0xffffffff for b in range(4)]
    for code in range(0, 13):
        assert L.mz_strerror(code)


def test_reference_messages(ffi):
    L = ffi.lib()
    assert b"must be odd to guarantee canonicality" in L.mz_strerror(4)     # src/canonical.rs:15
    assert b"Try splitting the input into 4GB chunks" in L.mz_strerror(3)   # src/sliding_min.rs:98
    assert b"Open syncmers require odd window size" in L.mz_strerror(5)     # src/syncmers.rs:27


def test_no_gpu_fails_loudly(ffi):
    """Without a CUDA device the library must refuse, never fall back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = ffi.lib()
    h = C.c_void_p()
    rc = L.mz_ctx_create(None, 0, C.byref(h))
    assert rc in (10, 11) and not h


def test_host_mirror_api_surface(sm):
    for name in ("minimizer_positions", "canonical_minimizer_positions", "minimizers",
                 "canonical_minimizers", "closed_syncmers", "canonical_closed_syncmers",
                 "open_syncmers", "canonical_open_syncmers", "canonical_syncmers"):
        assert callable(getattr(sm, name))
    b = sm.canonical_minimizers(5, 7).hasher(sm.MulHasher(5))
    sk = sm.U32Vec()
    b2 = b.super_kmers(sk)
    with pytest.raises(TypeError):
        b2.hasher(sm.NtHasher(5))            # src/lib.rs:323-338: hasher() only before super_kmers()
    with pytest.raises(TypeError):
        sm.closed_syncmers(5, 7).super_kmers(sk)   # src/lib.rs:339
    s = sm.PackedSeqVec.from_ascii(b"ACGTGCTCAGAGACTCAGAGGA")
    assert [s.as_slice().get(i) for i in range(4)] == [0, 1, 3, 2]
    rc = s.as_slice().to_revcomp()
    assert [rc.as_slice().get(i) for i in range(3)] == [2, 1, 1]   # T C C = revcomp of ...GGA


def test_header_is_plain_c_and_example_links(tmp_path):
    """include/mz_b200.h must be usable from C (the FFI boundary): the C99 example compiles with
    -pedantic and links against libmzb200.so (running it needs a GPU: tests/test_gpu_parity.py)."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "simd-minimizers_b200")
    if not os.path.exists(os.path.join(lib, "libmzb200.so")):
        import __graft_entry__ as g

        g.build()
    exe = str(tmp_path / "minimal")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic",
                           "-I", os.path.join(root, "include"), os.path.join(root, "examples", "minimal.c"),
                           "-L", lib, "-lmzb200", "-Wl,-rpath," + lib, "-o", exe])
    assert os.path.exists(exe)


def test_launch_geometry_invariants(tmp_path):
    """tests/plan_test.cu: host-only checks of plan_fast for every window length 1..300, all
    builders and input sizes from 1 to 2^32 windows (shared-memory limit, descriptor field widths,
    whole-iteration segments, sub-window / ring rules of the long-window instances)."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "plan_test")
    subprocess.check_call([os.environ.get("NVCC", "nvcc"), "-std=c++17", "-O1", "-gencode",
                           "arch=compute_100a,code=sm_100a", "-o", exe,
                           os.path.join(root, "tests", "plan_test.cu")])
    env = {k: v for k, v in os.environ.items() if not k.startswith("MZ_")}
    assert "plan ok" in subprocess.check_output([exe], env=env).decode()
