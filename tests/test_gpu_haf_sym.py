"""Symmetric-half hafnian kernel (thewalrus_b200/csrc/hafnian_sym.cu; even n in [36, 64] (whole-tile shapes, one-pair tails, and the sizes that run zero-padded in them), no loops, large ranges) against
(i) the C oracle's restatement of the reference sum (thewalrus/_hafnian.py:416-467) on the same subset ranges,
(ii) the row-panel kernel on the same ranges (WB200_HAF_SYM=0), (iii) exact closed forms of complete hafnians.
The complete n = 50 / 56 goldens of tests/test_gpu_fullsize.py run through this kernel too (it is the default for these sizes).

Tolerance 1e-10 relative (north_star); measured gaps are printed."""
import math
import os

import numpy as np
import pytest
from conftest import rel

import thewalrus_b200 as wb
from oracle import c_oracle as co
from thewalrus_b200 import _engine

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _input(n, seed, real=False):
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((n, n)) + (0 if real else 1j * rng.standard_normal((n, n)))
    A = (G + G.T).astype(np.complex128)
    x = co.matched_order(A)
    return np.ascontiguousarray(A[np.ix_(x, x)])


def _range(Ax, j0, j1, sym):
    old = os.environ.get("WB200_HAF_SYM")
    os.environ["WB200_HAF_SYM"] = str(int(sym))
    try:
        return _engine.combine4([_engine.hafnian_range(Ax, None, j0, j1)])
    finally:
        if old is None:
            del os.environ["WB200_HAF_SYM"]
        else:
            os.environ["WB200_HAF_SYM"] = old


@pytest.mark.parametrize("n", [36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56, 58, 60, 62, 64])
@pytest.mark.parametrize("real", [False, True])
def test_sym_kernel_ranges_match_oracle_and_row_panel_kernel(n, real):
    """Aligned, ragged and offset ranges (the kernel works on groups of four subsets; a range need not be a multiple)."""
    Ax = _input(n, 4800 + n + real, real)
    for j0, j1 in ((0, 8192), (12345, 12345 + 8195), ((1 << (n // 2 - 1)) - 6001, 1 << (n // 2 - 1))):
        want = co.hafnian_range(Ax, j0, j1)
        s, p = _range(Ax, j0, j1, True), _range(Ax, j0, j1, False)
        print(f"\n[haf_sym] n={n} real={real} [{j0}, {j1}): sym vs oracle {rel(s, want):.2e}, panel vs oracle {rel(p, want):.2e}, "
              f"sym vs panel {rel(s, p):.2e}")
        assert rel(s, want) <= TOL and rel(p, want) <= TOL and rel(s, p) <= TOL


def test_sym_kernel_range_additivity():
    """[0, b) = [0, a) + [a, b) for an unaligned split point: no subset is dropped or counted twice at a shard boundary."""
    Ax = _input(50, 99)
    a, b = 70001, 150003
    whole = _range(Ax, 0, b, True)
    parts = _range(Ax, 0, a, True) + _range(Ax, a, b, True)
    print(f"\n[haf_sym] additivity: {rel(parts, whole):.2e}")
    assert rel(parts, whole) <= 1e-12


def test_complete_hafnian_of_all_ones_48_is_double_factorial():
    """haf(J_48) = 47!! (number of perfect matchings of K_48): all 2^23 subsets through the symmetric-half kernel."""
    want = float(math.prod(range(1, 48, 2)))
    got = wb.hafnian(np.ones((48, 48), dtype=np.complex128))
    print(f"\n[haf_sym] haf(J_48) = {got!r}  47!! = {want!r}  rel_err = {rel(got, want):.2e}")
    assert rel(got, want) <= TOL


def test_complete_hafnian_of_all_ones_46_is_double_factorial():
    """haf(J_46) = 45!!: a size that runs zero-padded in the n = 48 shape, all 2^22 subsets."""
    want = float(math.prod(range(1, 46, 2)))
    got = wb.hafnian(np.ones((46, 46), dtype=np.complex128))
    print(f"\n[haf_sym] haf(J_46) = {got!r}  45!! = {want!r}  rel_err = {rel(got, want):.2e}")
    assert rel(got, want) <= TOL


def test_complete_hafnian_48_block_factorisation():
    """haf(A1 (+) A2) = haf(A1) haf(A2) with 24 + 24 vertices, shuffled: complete n = 48 run against two n = 24 runs
    (row-panel kernel) and, for those, the C oracle."""
    rng = np.random.default_rng(4824)
    blocks = []
    for _ in range(2):
        G = rng.standard_normal((24, 24)) + 1j * rng.standard_normal((24, 24))
        blocks.append(G + G.T)
    A = np.zeros((48, 48), dtype=np.complex128)
    A[:24, :24], A[24:, 24:] = blocks
    perm = rng.permutation(48)
    A = A[np.ix_(perm, perm)]
    h1, h2 = wb.hafnian(blocks[0]), wb.hafnian(blocks[1])
    assert rel(h1, co.hafnian(blocks[0])) <= TOL
    got = wb.hafnian(A)
    print(f"\n[haf_sym] n=48 direct sum: rel_err = {rel(got, h1 * h2):.2e}")
    assert rel(got, h1 * h2) <= 1e-9      # the product of two sums of 2^11 cancelling terms each; measured ~1e-12


@pytest.mark.parametrize("n", [50, 56])
def test_row_panel_kernel_still_serves_when_asked(n):
    """WB200_HAF_SYM=0 keeps the round-1 kernel: both kernels on 2^20 subsets of the same input, 1e-10; =4 the one-team shape."""
    Ax = _input(n, 5000 + n)
    s = _range(Ax, 0, 1 << 20, True)
    p = _range(Ax, 0, 1 << 20, False)
    print(f"\n[haf_sym] n={n}, 2^20 subsets: sym vs panel {rel(s, p):.2e}")
    assert rel(s, p) <= TOL
    if n == 50:
        one = _range(Ax, 0, 1 << 20, 4)
        print(f"[haf_sym] n={n}: one team of 12 warps vs two teams of 6: {rel(one, s):.2e}")
        assert rel(one, s) <= TOL
