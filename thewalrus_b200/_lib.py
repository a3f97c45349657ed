"""ctypes binding of libwalrus_b200.so (the C ABI declared in include/walrus_b200.h).

There is deliberately no CPU fallback: if the shared library is missing, or no CUDA device is
visible, every compute call raises ``RuntimeError``.
"""
import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwalrus_b200.so")

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int32_p = ctypes.POINTER(ctypes.c_int32)
_c_int64_p = ctypes.POINTER(ctypes.c_int64)
_c_uint64_p = ctypes.POINTER(ctypes.c_uint64)
_u64 = ctypes.c_uint64
_vp = ctypes.c_void_p

# symbol -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "wb200_last_error": (ctypes.c_char_p, []),
    "wb200_version": (ctypes.c_int, []),
    "wb200_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "wb200_release_scratch": (ctypes.c_int, []),
    "wb200_fp64_peak": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _c_double_p]),
    "wb200_hafnian_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int]),
    "wb200_hafnian_dev": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _u64, _u64, _vp, _vp, ctypes.c_size_t, _vp]),
    "wb200_hafnian_host": (ctypes.c_int, [ctypes.c_int, _c_double_p, _c_double_p, ctypes.c_int, _u64, _u64,
                                          _c_double_p, _c_double_p]),
    "wb200_lhaf_general_steps": (ctypes.c_int, [_c_int32_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_uint64_p]),
    "wb200_lhaf_general_host": (ctypes.c_int, [ctypes.c_int, _c_double_p, _c_double_p, _c_double_p, _c_double_p,
                                               ctypes.c_int, _c_int32_p, ctypes.c_int, _u64, _u64, _c_double_p,
                                               _c_double_p]),
    "wb200_lhaf_general_dev": (ctypes.c_int, [_vp, _vp, _vp, _c_double_p, ctypes.c_int, _c_int32_p, ctypes.c_int, _u64, _u64,
                                              _vp, _vp]),
    "wb200_lhaf_matrices_dev": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, ctypes.c_int, _vp, ctypes.c_int, _vp,
                                               ctypes.c_int64, ctypes.c_int, _vp, _vp]),
    "wb200_lhaf_batch_gamma_dev": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, _c_int32_p, ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_int, _u64, _u64, _vp, ctypes.c_int, _vp]),
    "wb200_mtl_dev": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _u64, _u64, _vp, _vp]),
    "wb200_brs_dev": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, _u64, _u64, _vp, _vp]),
    "wb200_lhaf_patterns_host": (ctypes.c_int, [ctypes.c_int, _c_double_p, _c_double_p, ctypes.c_int, _c_int32_p,
                                                ctypes.c_int64, ctypes.c_int, _c_double_p, _c_double_p]),
    "wb200_lhaf_patterns_multi_host": (ctypes.c_int, [ctypes.c_int, _c_double_p, _c_double_p, ctypes.c_int, _c_int32_p,
                                                      ctypes.c_int, _c_int32_p, ctypes.c_int64, ctypes.c_int,
                                                      _c_double_p, _c_double_p]),
    "wb200_lhaf_matrices_host": (ctypes.c_int, [ctypes.c_int, _c_double_p, ctypes.c_int, _c_int32_p, _c_double_p,
                                                ctypes.c_int, _c_int32_p, ctypes.c_int, _c_int32_p, ctypes.c_int64,
                                                ctypes.c_int, _c_double_p, _c_double_p]),
    "wb200_lhaf_batch_steps": (ctypes.c_int, [_c_int32_p, ctypes.c_int, _c_uint64_p]),
    "wb200_lhaf_batch_host": (ctypes.c_int, [ctypes.c_int, _c_double_p, _c_double_p, ctypes.c_int, _c_int32_p,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int, _u64, _u64, _c_double_p,
                                             ctypes.c_int, _c_double_p]),
    "wb200_lhaf_batch_gamma_host": (ctypes.c_int, [ctypes.c_int, _c_double_p, _c_double_p, ctypes.c_int, ctypes.c_int,
                                                   _c_int32_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _u64, _u64,
                                                   _c_double_p, ctypes.c_int, _c_double_p]),
    "wb200_hafnian_chains_host": (ctypes.c_int, [ctypes.c_int, _c_double_p, _c_double_p, _c_double_p, _c_double_p, ctypes.c_int,
                                                 ctypes.c_int64, ctypes.c_int, _c_int32_p, _c_double_p]),
    "wb200_mtl_host": (ctypes.c_int, [ctypes.c_int, _c_double_p, _c_double_p, ctypes.c_int, _u64, _u64, _c_double_p,
                                      _c_double_p]),
    "wb200_perm_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int]),
    "wb200_perm_dev": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _u64, _u64, _vp, _vp, ctypes.c_size_t, _vp]),
    "wb200_perm_host": (ctypes.c_int, [ctypes.c_int, _c_double_p, ctypes.c_int, ctypes.c_int, _u64, _u64,
                                       _c_double_p, _c_double_p]),
    "wb200_perm_f64_host": (ctypes.c_int, [ctypes.c_int, _c_double_p, ctypes.c_int, ctypes.c_int, _u64, _u64,
                                           _c_double_p, _c_double_p]),
    "wb200_perm_int64_host": (ctypes.c_int, [ctypes.c_int, _c_int64_p, ctypes.c_int, ctypes.c_int, _u64, _u64,
                                             _c_int64_p, _c_double_p]),
    "wb200_brs_host": (ctypes.c_int, [ctypes.c_int, _c_double_p, _c_double_p, ctypes.c_int, ctypes.c_int, _u64, _u64,
                                      _c_double_p, _c_double_p]),
    "wb200_tor_num_prefixes": (ctypes.c_int, [ctypes.c_int, _c_uint64_p]),
    "wb200_tor_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int]),
    "wb200_tor_dev": (ctypes.c_int, [_vp, ctypes.c_int, _u64, _u64, _vp, _vp, ctypes.c_size_t, _vp]),
    "wb200_tor_host": (ctypes.c_int, [ctypes.c_int, _c_double_p, ctypes.c_int, _u64, _u64, _c_double_p,
                                      _c_double_p]),
    "wb200_tor_f64_dev": (ctypes.c_int, [_vp, ctypes.c_int, _u64, _u64, _vp, _vp, ctypes.c_size_t, _vp]),
    "wb200_tor_f64_host": (ctypes.c_int, [ctypes.c_int, _c_double_p, ctypes.c_int, _u64, _u64, _c_double_p,
                                          _c_double_p]),
    "wb200_ltor_dev": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _u64, _u64, _vp, _vp, ctypes.c_size_t, _vp]),
    "wb200_ltor_host": (ctypes.c_int, [ctypes.c_int, _c_double_p, _c_double_p, ctypes.c_int, _u64, _u64, _c_double_p,
                                       _c_double_p]),
}

_lib = None
_lock = threading.Lock()


def load():
    """Load the shared library (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"walrus_b200 CUDA library not built: {LIB_PATH} is missing. "
                    "Run `python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)."
                )
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().wb200_last_error().decode("utf-8", "replace")
        if rc == -3:
            raise NotImplementedError(f"{what}: {msg}")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def as_c128(a):
    """C-contiguous complex128 copy viewed as interleaved float64."""
    a = np.ascontiguousarray(a, dtype=np.complex128)
    return a, a.view(np.float64).ctypes.data_as(_c_double_p)


def dptr(arr):
    return arr.ctypes.data_as(_c_double_p)
