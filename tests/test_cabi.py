"""The C-ABI shared library loads and exports every symbol include/walrus_b200.h declares (no compute
calls here: this file runs on the CPU-only box)."""
import ctypes
import os
import re

from conftest import ROOT

from thewalrus_b200 import _lib


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "walrus_b200.h")) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wb200_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/walrus_b200.h but not exported"


def test_ctypes_table_covers_header():
    assert set(_declared_symbols()) == set(_lib.SIGNATURES)


def test_header_compiles_as_c():
    import subprocess
    import tempfile

    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.c")
        with open(src, "w") as fh:
            fh.write('#include "walrus_b200.h"\nint main(void){return wb200_version() > 0 ? 0 : 1;}\n')
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", src,
                               "-o", os.path.join(td, "t.o")])


def test_argument_checks_do_not_need_a_gpu():
    lib = _lib.load()
    assert lib.wb200_version() >= 100
    assert lib.wb200_hafnian_workspace_bytes(7) == 0      # odd n rejected
    assert lib.wb200_hafnian_workspace_bytes(66) == 0     # beyond the kernel limit
    assert lib.wb200_hafnian_workspace_bytes(50) > 0
    steps = ctypes.c_uint64(0)
    reps = (ctypes.c_int32 * 3)(3, 2, 1)
    assert lib.wb200_lhaf_general_steps(reps, 3, 1, 0, ctypes.byref(steps)) == 0 and steps.value == 12
    assert lib.wb200_lhaf_general_steps(reps, 3, 0, 0, ctypes.byref(steps)) == 0 and steps.value == 24
    cnt = ctypes.c_uint64(0)
    assert lib.wb200_tor_num_prefixes(24, ctypes.byref(cnt)) == 0 and cnt.value == 1 << 15
    assert lib.wb200_tor_num_prefixes(40, ctypes.byref(cnt)) == -3
    assert b"modes" in lib.wb200_last_error()


def test_argument_checks_of_the_newer_entry_points():
    import numpy as np

    lib = _lib.load()
    O = np.zeros(2 * 8 * 8)
    out = np.zeros(4)
    assert lib.wb200_ltor_host(0, _lib.dptr(O), None, 4, 0, 1, _lib.dptr(out), None) == -1      # gamma is required
    assert b"gamma" in lib.wb200_last_error()
    A = np.zeros(2 * 2 * 3 * 3)
    rpt = np.ones((2, 3), dtype=np.int32)
    prpt = rpt.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
    res = np.zeros(4)
    # two matrices but no index / index out of range / table of loop vectors without an index
    assert lib.wb200_lhaf_matrices_host(0, _lib.dptr(A), 2, None, None, 0, None, 3, prpt, 2, 1, _lib.dptr(res), None) == -1
    bad = np.array([0, 2], dtype=np.int32)
    assert lib.wb200_lhaf_matrices_host(0, _lib.dptr(A), 2, bad.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), None, 0,
                                        None, 3, prpt, 2, 1, _lib.dptr(res), None) == -1
    assert b"A_index" in lib.wb200_last_error()
    g = np.zeros(2 * 2 * 3)
    assert lib.wb200_lhaf_patterns_multi_host(0, _lib.dptr(A), _lib.dptr(g), 2, None, 3, prpt, 2, 1, _lib.dptr(res),
                                              None) == -1
    assert lib.wb200_lhaf_patterns_multi_host(0, _lib.dptr(A), None, 0, None, 65, prpt, 2, 1, _lib.dptr(res), None) == -3
    assert lib.wb200_release_scratch() == 0
