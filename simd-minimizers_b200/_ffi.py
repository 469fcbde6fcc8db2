"""ctypes binding of libmzb200.so -- the C ABI declared in include/mz_b200.h.

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is
no CPU fallback: if the library is missing, or no CUDA device is present, every compute call
raises.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# MZ_B200_LIB: another build of the same library (A/B experiments with compile-time variants)
LIB_PATH = os.environ.get("MZ_B200_LIB") or os.path.join(_PKG, "libmzb200.so")

MZ_OK = 0
ERR_NAMES = {
    1: "MZ_ERR_BAD_ARG", 2: "MZ_ERR_W_RANGE", 3: "MZ_ERR_TOO_LONG", 4: "MZ_ERR_EVEN_L",
    5: "MZ_ERR_OPEN_EVEN_W", 6: "MZ_ERR_NOT_CANONICAL", 7: "MZ_ERR_VALUE_WIDTH",
    8: "MZ_ERR_CAPACITY", 9: "MZ_ERR_UNSUPPORTED", 10: "MZ_ERR_NO_DEVICE", 11: "MZ_ERR_CUDA",
    12: "MZ_ERR_NOMEM",
}
MZ_ERR_CAPACITY = 8
MODE_MINIMIZER, MODE_CLOSED_SYNCMER, MODE_OPEN_SYNCMER = 0, 1, 2

# every symbol include/mz_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "mz_abi_version", "mz_strerror", "mz_last_error", "mz_device_count", "mz_params_nthash",
    "mz_params_mulhash", "mz_params_set_nthash", "mz_params_set_mulhash", "mz_params_validate",
    "mz_ctx_create", "mz_ctx_destroy", "mz_ctx_device_count", "mz_host_alloc", "mz_host_free",
    "mz_run", "mz_run_device", "mz_run_batch", "mz_pack_ascii", "mz_run_ascii", "mz_last_timing",
    "mz_run_skip_ambiguous", "mz_run_device_skip_ambiguous", "mz_pack_ascii_n",
    "mz_run_ascii_skip_ambiguous", "mz_params_set_tables", "mz_values", "mz_pcie_probe", "mz_alu_probe", "mz_run_bucket_stats",
]


class MzParams(C.Structure):
    _fields_ = [("k", C.c_uint32), ("w", C.c_uint32), ("mode", C.c_uint32),
                ("strand_tiebreak", C.c_uint32), ("hash_canonical", C.c_uint32),
                ("rot", C.c_uint32), ("f", C.c_uint32 * 4), ("c", C.c_uint32 * 4),
                ("want_sk", C.c_uint32), ("value_bits", C.c_uint32), ("reserved", C.c_uint32 * 2)]


class MzOut(C.Structure):
    _fields_ = [("pos", C.c_void_p), ("sk", C.c_void_p), ("val", C.c_void_p),
                ("capacity", C.c_uint64), ("count", C.c_uint64)]


class MzTiming(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("kernel_ms", C.c_float), ("d2h_ms", C.c_float),
                ("total_ms", C.c_float), ("kernel_launches", C.c_uint32), ("reserved", C.c_uint32)]


class MzPcieResult(C.Structure):
    _fields_ = [("h2d_gbs", C.c_double), ("d2h_gbs", C.c_double), ("bidir_h2d_gbs", C.c_double),
                ("bidir_d2h_gbs", C.c_double), ("n_devices", C.c_uint32), ("reserved", C.c_uint32)]


class MzAluResult(C.Structure):
    _fields_ = [("lane_ops_per_s", C.c_double), ("ms", C.c_float), ("sm_count", C.c_uint32)]


class MzError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load libmzb200.so (raises if it has not been built -- there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a). simd-minimizers_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.mz_abi_version.restype = C.c_uint32
    L.mz_strerror.restype = C.c_char_p
    L.mz_strerror.argtypes = [C.c_int]
    L.mz_last_error.restype = C.c_char_p
    L.mz_device_count.argtypes = [C.POINTER(C.c_int)]
    for fn in (L.mz_params_nthash, L.mz_params_mulhash):
        fn.argtypes = [C.POINTER(MzParams), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    for fn in (L.mz_params_set_nthash, L.mz_params_set_mulhash):
        fn.argtypes = [C.POINTER(MzParams), C.c_uint32]
    L.mz_params_validate.argtypes = [C.POINTER(MzParams), C.c_uint64]
    L.mz_ctx_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    L.mz_ctx_destroy.argtypes = [vp]
    L.mz_ctx_destroy.restype = None
    L.mz_ctx_device_count.argtypes = [vp]
    L.mz_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.mz_host_free.argtypes = [vp]
    L.mz_host_free.restype = None
    L.mz_run.argtypes = [vp, C.POINTER(MzParams), vp, C.c_uint64, C.c_uint64, C.POINTER(MzOut)]
    L.mz_run_device.argtypes = [vp, C.c_int, C.POINTER(MzParams), vp, C.c_uint64, C.c_uint64,
                                C.c_uint64, C.c_uint64, C.POINTER(MzOut)]
    L.mz_run_batch.argtypes = [vp, C.POINTER(MzParams), vp, C.c_uint64, C.c_uint64, vp, vp,
                               C.c_uint64, C.c_uint32, vp, C.POINTER(MzOut)]
    L.mz_pack_ascii.argtypes = [vp, C.c_char_p, C.c_uint64, vp]
    L.mz_run_ascii.argtypes = [vp, C.POINTER(MzParams), C.c_char_p, C.c_uint64, C.POINTER(MzOut)]
    L.mz_last_timing.argtypes = [vp, C.POINTER(MzTiming)]
    L.mz_run_skip_ambiguous.argtypes = [vp, C.POINTER(MzParams), vp, C.c_uint64, C.c_uint64, vp,
                                        C.c_uint64, C.POINTER(MzOut)]
    L.mz_run_device_skip_ambiguous.argtypes = [vp, C.c_int, C.POINTER(MzParams), vp, C.c_uint64,
                                               C.c_uint64, vp, C.c_uint64, C.c_uint64, C.c_uint64,
                                               C.POINTER(MzOut)]
    L.mz_pack_ascii_n.argtypes = [vp, C.c_char_p, C.c_uint64, vp, vp]
    L.mz_run_ascii_skip_ambiguous.argtypes = [vp, C.POINTER(MzParams), C.c_char_p, C.c_uint64,
                                              C.POINTER(MzOut)]
    L.mz_params_set_tables.argtypes = [C.POINTER(MzParams), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                       C.c_uint32, C.c_uint32]
    L.mz_values.argtypes = [vp, C.POINTER(MzParams), vp, C.c_uint64, C.c_uint64, vp, C.c_uint64,
                            C.c_uint32, vp]
    if not os.environ.get("MZ_B200_LIB") or hasattr(L, "mz_run_bucket_stats"):  # (older A/B variants lack it)
        L.mz_run_bucket_stats.argtypes = [vp, C.POINTER(MzParams), vp, C.c_uint64, C.c_uint64, C.c_uint32, vp, vp,
                                          C.POINTER(C.c_uint64)]
    L.mz_alu_probe.argtypes = [vp, C.c_int, C.POINTER(MzAluResult)]
    L.mz_pcie_probe.argtypes = [vp, C.c_uint64, C.c_uint32, C.POINTER(MzPcieResult)]
    _lib = L
    return L


def check(code: int):
    if code != MZ_OK:
        L = lib()
        msg = L.mz_strerror(code).decode()
        if code == 11:
            msg += ": " + L.mz_last_error().decode()
        raise MzError(code, msg)
