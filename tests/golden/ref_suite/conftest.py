"""Run the REFERENCE's own test files against thewalrus_b200 ("next round" item 2).

The `test_*.py` files in this directory are verbatim copies of /root/reference/thewalrus/tests/{test_hafnian,
test_hafnian_repeated, test_permanent, test_torontonian, test_montrealer}.py — test fixtures (known answers and
wrapper-validation cases), not product code; nothing under thewalrus_b200/ imports them.  This conftest

  * installs thewalrus_b200 as ``sys.modules["thewalrus"]`` (+ the submodule names the tests import:
    ``_hafnian``, ``_permanent``, ``_torontonian``, ``quantum``, ``random``, ``symplectic``, ``reference``),
  * provides the fixtures of the reference's tests/conftest.py (:23-61: ``tol``, ``hbar``, ``dtype``,
    ``random_matrix``), same seed,
  * marks every test here ``gpu`` (they call the CUDA kernels), and
  * skips BY NAME the tests of functions SURVEY 2 / 8 put outside the hot path, each with the row cited.

``WB200_REFSUITE_CPU=1`` answers the kernel calls with the CPU oracle instead (tests/oracle_engine.py) so the host
logic of the shim can be checked on a box without a GPU (tests/test_ref_suite_cpu.py drives that).
"""
import os
import sys
import types

import numpy as np
import pytest

import thewalrus_b200 as wb
from thewalrus_b200 import _hafnian, _permanent, _torontonian, moments, quantum

CPU = os.environ.get("WB200_REFSUITE_CPU") == "1"


# ---- helpers the reference tests import from modules outside the hot path (small host-side utilities, written here) -------------
def random_covariance(N, hbar=2, pure=False, block_diag=False):
    """A valid random N-mode covariance matrix: S diag(nu, nu) S^T (hbar/2) with S from a Bloch-Messiah product of two
    Haar interferometers around single-mode squeezers (role of thewalrus/random.py:51-75)."""
    del block_diag

    def haar(n):
        Z = (np.random.randn(n, n) + 1j * np.random.randn(n, n)) / np.sqrt(2)
        Q, R = np.linalg.qr(Z)
        return Q * (np.diag(R) / np.abs(np.diag(R)))

    def passive(U):
        return np.block([[U.real, -U.imag], [U.imag, U.real]])

    r = np.random.rand(N)
    S = passive(haar(N)) @ np.diag(np.concatenate([np.exp(-r), np.exp(r)])) @ passive(haar(N))
    nu = np.ones(N) if pure else 1.0 + np.random.rand(N)
    return (hbar / 2) * S @ np.diag(np.concatenate([nu, nu])) @ S.T


def two_mode_squeezing(r, phi=0.0):
    """Symplectic matrix of a two-mode squeezer in xxpp order (role of thewalrus/symplectic.py two_mode_squeezing)."""
    cp, sp, ch, sh = np.cos(phi), np.sin(phi), np.cosh(r), np.sinh(r)
    return np.array([[ch, cp * sh, 0, sp * sh], [cp * sh, ch, sp * sh, 0], [0, sp * sh, ch, -cp * sh],
                     [sp * sh, 0, -cp * sh, ch]])


def _outside(name, row):
    def stub(*a, **k):
        raise NotImplementedError(f"{name} is outside the hot path ({row})")

    stub.__name__ = name
    return stub


def _install():
    tw = types.ModuleType("thewalrus")
    for k in dir(wb):
        if not k.startswith("__"):
            setattr(tw, k, getattr(wb, k))
    tw.__path__ = []     # a package: the tests import thewalrus._hafnian etc.
    tw.version = wb.version
    tw.hafnian_sparse = _outside("hafnian_sparse", "SURVEY 2: sparse/banded hafnians, out of scope")
    tw.hafnian_banded = _outside("hafnian_banded", "SURVEY 2: sparse/banded hafnians, out of scope")

    m_haf = types.ModuleType("thewalrus._hafnian")
    for k in dir(_hafnian):
        if not k.startswith("__"):
            setattr(m_haf, k, getattr(_hafnian, k))
    m_haf.bandwidth = _outside("bandwidth", "SURVEY 2: banded hafnian helper, out of scope")
    if not hasattr(m_haf, "recursive_hafnian"):
        m_haf.recursive_hafnian = lambda A: _hafnian.hafnian(np.asarray(A), method="recursive")

    m_perm = types.ModuleType("thewalrus._permanent")
    for k in dir(_permanent):
        if not k.startswith("__"):
            setattr(m_perm, k, getattr(_permanent, k))
    m_tor = types.ModuleType("thewalrus._torontonian")
    for k in dir(_torontonian):
        if not k.startswith("__"):
            setattr(m_tor, k, getattr(_torontonian, k))

    m_q = types.ModuleType("thewalrus.quantum")
    for mod in (quantum, moments):
        for k in dir(mod):
            if not k.startswith("_"):
                setattr(m_q, k, getattr(mod, k))
    m_rand = types.ModuleType("thewalrus.random")
    m_rand.random_covariance = random_covariance
    m_symp = types.ModuleType("thewalrus.symplectic")
    m_symp.two_mode_squeezing = two_mode_squeezing
    m_ref = types.ModuleType("thewalrus.reference")    # symbolic montrealer helpers: thewalrus/reference.py, SURVEY 2 "reference" row
    for name in ("mapper", "rspm", "rpmp", "mtl"):
        setattr(m_ref, name, _outside(f"reference.{name}", "SURVEY 2: thewalrus.reference symbolic helpers, out of scope"))

    for name, mod in (("thewalrus", tw), ("thewalrus._hafnian", m_haf), ("thewalrus._permanent", m_perm),
                      ("thewalrus._torontonian", m_tor), ("thewalrus.quantum", m_q), ("thewalrus.random", m_rand),
                      ("thewalrus.symplectic", m_symp), ("thewalrus.reference", m_ref)):
        sys.modules[name] = mod
        if "." in name:
            setattr(tw, name.split(".", 1)[1], mod)


_install()
np.random.seed(137)      # thewalrus/tests/conftest.py:20

# tests of functions outside the hot path: skipped by NAME, with the SURVEY row that excludes the function
OUT_OF_SCOPE = {
    "test_hafnian.py::test_valid_output": "compares with hafnian_sparse (SURVEY 2: sparse/banded hafnians — not on the north-star path)",
    "test_hafnian.py::test_hafnian_banded": "hafnian_banded (SURVEY 2: sparse/banded hafnians)",
    "test_hafnian.py::test_bandwidth": "bandwidth helper of hafnian_banded (SURVEY 2)",
    "test_hafnian.py::test_bandwidth_zero": "bandwidth helper of hafnian_banded (SURVEY 2)",
    "test_montrealer.py::test_size_of_rpmp": "thewalrus.reference symbolic matchings (SURVEY 2: reference.py, not the path)",
    "test_montrealer.py::test_size_of_rspm": "thewalrus.reference symbolic matchings (SURVEY 2)",
    "test_montrealer.py::test_rpmp_alternating_walk": "thewalrus.reference symbolic matchings (SURVEY 2)",
    "test_montrealer.py::test_mtl_functions_agree": "compares with thewalrus.reference.mtl, the symbolic montrealer (SURVEY 2)",
    "test_montrealer.py::test_lmtl_functions_agree": "compares with thewalrus.reference.mtl, the symbolic montrealer (SURVEY 2)",
    "test_montrealer.py::test_mtl_lmtl_reference_agree": "tests thewalrus.reference.mtl only (SURVEY 2)",
    "test_montrealer.py::test_mapper_hard_coded": "tests thewalrus.reference.mapper only (SURVEY 2)",
}


def pytest_collection_modifyitems(config, items):
    here = os.path.dirname(os.path.abspath(__file__))
    for item in items:
        if not str(item.fspath).startswith(here):
            continue
        if not CPU:
            item.add_marker(pytest.mark.gpu)
        base = os.path.basename(str(item.fspath))
        fn = item.originalname if hasattr(item, "originalname") and item.originalname else item.name
        why = OUT_OF_SCOPE.get(f"{base}::{fn}")
        if why:
            item.add_marker(pytest.mark.skip(reason="out of scope: " + why))


@pytest.fixture(autouse=True)
def _kernel(monkeypatch):
    if CPU:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
        import oracle_engine

        oracle_engine.install(monkeypatch)
    yield


# ---- fixtures of thewalrus/tests/conftest.py:23-61 ------------------------------------------------------------------------------
@pytest.fixture(scope="session")
def tol():
    return 1e-3


@pytest.fixture(params=[0.5, 1, 2])
def hbar(request):
    return request.param


@pytest.fixture(params=[np.complex128, np.float64, np.int64])
def dtype(request):
    return request.param


@pytest.fixture
def random_matrix(dtype):
    """Random symmetric n x n matrix of type ``dtype`` (same recipe as the reference fixture)."""

    def _wrapper(n, fill_factor=1.0):
        A = np.complex128(np.random.random([n, n]))
        A += 1j * np.random.random([n, n])
        A *= np.random.binomial(1, p=fill_factor, size=(n, n))
        A += A.T
        if not np.issubdtype(dtype, np.complexfloating):
            A = A.real
        return A.astype(dtype)

    return _wrapper
