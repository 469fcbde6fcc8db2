"""CPU test of the host half of the transfer codec (simd-minimizers_b200/csrc/mz_host_codec.cpp):
the AVX2 / scalar delta decoder against numpy, for every output misalignment (a chunk's output
starts wherever the chunks before it ended), ragged last blocks and block sub-ranges."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_delta_decoder_matches_numpy(tmp_path):
    so = str(tmp_path / "libcodec.so")
    shim = tmp_path / "shim.cpp"
    shim.write_text(
        '#include <stdint.h>\n'
        'namespace mz { void delta_decode_blocks(const int8_t*, const uint32_t*, uint64_t, uint64_t, uint64_t, uint32_t*); }\n'
        'extern "C" void decode(const int8_t* d, const uint32_t* b, uint64_t n, uint64_t b0, uint64_t b1, uint32_t* o) '
        '{ mz::delta_decode_blocks(d, b, n, b0, b1, o); }\n')
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, str(shim),
                           os.path.join(ROOT, "simd-minimizers_b200", "csrc", "mz_host_codec.cpp")])
    L = C.CDLL(so)
    L.decode.argtypes = [C.c_void_p] * 2 + [C.c_uint64] * 3 + [C.c_void_p]
    rng = np.random.default_rng(3)
    for n in (1, 7, 255, 256, 257, 511, 1000, 4096 + 13, 100_003):
        vals = (np.cumsum(rng.integers(-127, 128, size=n)).astype(np.int64) + (1 << 31)).astype(np.uint32)
        nblk = (n + 255) // 256
        delta = np.zeros(n + 8, dtype=np.int8)   # the decoder may load 8 bytes at a time
        delta[1:n] = (vals[1:].astype(np.int64) - vals[:-1].astype(np.int64)).astype(np.int8)
        delta[0:n:256] = 0
        base = vals[0:n:256].copy()
        for mis in range(8):                      # output address = 32-byte aligned + 4 * mis
            raw = np.zeros(n + 16 + 8, dtype=np.uint32)
            off = (-(raw.ctypes.data // 4) % 8 + mis) % 8 + 8
            out = raw[off:off + n]
            assert (out.ctypes.data % 32) == 4 * mis
            L.decode(delta.ctypes.data, base.ctypes.data, n, 0, nblk, out.ctypes.data)
            assert np.array_equal(out, vals), (n, mis)
            assert not raw[:off].any() and not raw[off + n:].any(), (n, mis)   # nothing written outside
            if nblk >= 3:                          # a sub-range of blocks (one decode thread's share)
                raw[:] = 0
                L.decode(delta.ctypes.data, base.ctypes.data, n, 1, nblk - 1, out.ctypes.data)
                assert np.array_equal(out[256:(nblk - 1) * 256], vals[256:(nblk - 1) * 256]), (n, mis)
                assert not out[:256].any() and not out[(nblk - 1) * 256:].any()
