"""Real-symmetric torontonian (wb200_tor_f64_host / _dev, tor_kernel<0, 256, double>): the subset tree of
thewalrus/_torontonian.py:157-247 in real arithmetic.  Checked against the long-double C oracle, against the complex
kernel on the same input (WB200_TOR_REAL=0), through prefix ranges, through the device-pointer twin on a non-default
stream, and on a closed form (diagonal O).  Tolerance 1e-10 relative."""
import ctypes
import os

import numpy as np
import pytest
from conftest import rel

import thewalrus_b200 as wb
from oracle import c_oracle as co
from thewalrus_b200 import _engine, _lib

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _real_O(N, seed):
    rng = np.random.default_rng(seed)
    B = rng.standard_normal((2 * N, 2 * N))
    H = B @ B.T
    return 0.9 * H / np.linalg.norm(H, 2)


def _complex_kernel(O):
    os.environ["WB200_TOR_REAL"] = "0"
    try:
        return wb.tor(O)
    finally:
        del os.environ["WB200_TOR_REAL"]


@pytest.mark.parametrize("N", [2, 3, 4, 7, 9, 10, 13, 16, 20])
def test_real_tor_vs_oracle_and_complex_kernel(N):
    O = _real_O(N, 7000 + N)
    got = wb.tor(O)
    assert isinstance(got, np.float64)                                    # the reference returns a real scalar for real input
    cplx = _complex_kernel(O)
    want = co.tor_recursive(O.astype(np.complex128), long_double=True) if N <= 16 else cplx
    print(f"\n[tor real] N={N}: vs oracle {rel(got, want):.2e}, vs complex kernel {rel(got, cplx):.2e}")
    assert rel(got, want) < TOL and rel(got, cplx) < TOL
    assert rel(wb.tor(O.astype(np.complex128)), want) < TOL               # complex dtype with zero imaginary part: same path
    total = _engine.tor_num_prefixes(N)
    if total >= 4:  # prefix ranges add up
        cuts = [0, 1, total // 2 + 1, total]
        s = sum(sum(_engine.tor_range(O, a, b)) for a, b in zip(cuts[:-1], cuts[1:]))
        assert rel(s, want) < TOL


def test_real_tor_diagonal_closed_form():
    """Diagonal O (uncorrelated thermal modes): every principal minor factorises, tor = prod_k (1 / (1 - o_k) - 1)."""
    N = 6
    t = np.tanh(0.7)
    O = np.zeros((2 * N, 2 * N))
    for k in range(N):
        O[k, k] = O[k + N, k + N] = t * t * (0.5 + 0.05 * k)
    want = np.prod([1.0 / (1.0 - O[k, k]) - 1.0 for k in range(N)])
    got = wb.tor(O)
    print(f"\n[tor real] diagonal O, N={N}: {rel(got, want):.2e}")
    assert rel(got, want) < TOL


def test_real_tor_dev_entry_on_a_stream_matches_host_entry():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    lib = _lib.load()
    N = 12
    O = np.ascontiguousarray(_real_O(N, 99))
    total = _engine.tor_num_prefixes(N)
    host = np.zeros(2)
    assert lib.wb200_tor_f64_host(0, _lib.dptr(O), N, 0, total, _lib.dptr(host), None) == 0
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream(dev)
    wsb = lib.wb200_tor_workspace_bytes(N)
    with torch.cuda.stream(st):
        dO = torch.from_numpy(O.reshape(-1).copy()).to(dev)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        out = torch.zeros(4, dtype=torch.float64, device=dev)
        rc = lib.wb200_tor_f64_dev(ctypes.c_void_p(dO.data_ptr()), N, 0, total, ctypes.c_void_p(out.data_ptr()),
                                   ctypes.c_void_p(ws.data_ptr()), wsb, ctypes.c_void_p(st.cuda_stream))
        assert rc == 0, lib.wb200_last_error()
    st.synchronize()
    o = out.cpu().numpy()
    assert o[0] == host[0] and o[1] == host[1]
