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
