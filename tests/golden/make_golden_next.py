"""Golden vectors for the SURVEY 8(f) "next" components, generated from the REFERENCE ITSELF
(/root/reference): loop_hafnian_batch_gamma, montrealer / loop montrealer, loop torontonian and
threshold_detection_prob, Bristolian.  Run once in the authoring container
(`python tests/golden/make_golden_next.py`); the output tests/golden/reference_outputs_next.json is committed.
"""
import json
import os
import sys
import types

import numpy as np

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")
_d = types.ModuleType("dask")
_d.delayed = lambda f, *a, **k: f
_d.compute = lambda *a, **k: a
sys.modules["dask"] = _d
sys.path.insert(0, "/root/reference")

import thewalrus  # noqa: E402
from thewalrus import ltor, lmtl, mtl  # noqa: E402
from thewalrus._permanent import brs, fock_threshold_prob, ubrs  # noqa: E402
from thewalrus._torontonian import numba_ltor, threshold_detection_prob  # noqa: E402
from thewalrus.loop_hafnian_batch_gamma import loop_hafnian_batch_gamma  # noqa: E402
from thewalrus.quantum import Qmat  # noqa: E402
from thewalrus.random import random_covariance, random_interferometer  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def enc(z):
    z = np.asarray(z)
    if np.iscomplexobj(z):
        return {"re": z.real.tolist(), "im": z.imag.tolist()}
    return {"re": z.astype(float).tolist(), "im": (0 * z.astype(float)).tolist()}


def main():
    out = {"batch_gamma": [], "mtl": [], "lmtl": [], "ltor": [], "threshold": [], "brs": [], "ubrs": [],
           "fock_threshold": []}
    rng = np.random.default_rng(20261018)
    # loop_hafnian_batch_gamma
    for trial in range(12):
        n = int(rng.integers(2, 6))
        G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        A = (G + G.T) / np.sqrt(n)
        n_D = int(rng.integers(1, 5))
        D = rng.standard_normal((n_D, n)) + 1j * rng.standard_normal((n_D, n))
        fixed = [int(r) for r in rng.integers(0, 3, n - 1)]
        cutoff = int(rng.integers(1, 6))
        for glynn in (True, False):
            out["batch_gamma"].append({"A": enc(A), "D": enc(D), "fixed": fixed, "cutoff": cutoff, "glynn": glynn,
                                       "value": enc(loop_hafnian_batch_gamma(A, D, np.array(fixed), cutoff, glynn=glynn))})
    # montrealer / loop montrealer: adjacency-like matrices from Gaussian states and generic complex symmetric ones
    for N in (1, 2, 3, 4, 5, 6, 7):
        cov = random_covariance(N, hbar=2)
        Q = Qmat(cov, hbar=2)
        O = np.eye(2 * N) - np.linalg.inv(Q)          # Hermitian, as used by the cumulant formulas
        zeta = 0.4 * (rng.standard_normal(2 * N) + 1j * rng.standard_normal(2 * N))
        G = rng.standard_normal((2 * N, 2 * N)) + 1j * rng.standard_normal((2 * N, 2 * N))
        S = (G + G.T) / (2 * N)
        for name, Mx in (("gaussian", O), ("symmetric", S)):
            out["mtl"].append({"N": N, "kind": name, "A": enc(Mx), "value": enc(mtl(Mx.astype(np.complex128)))})
            out["lmtl"].append({"N": N, "kind": name, "A": enc(Mx), "zeta": enc(zeta),
                                "value": enc(lmtl(Mx.astype(np.complex128), zeta))})
    # loop torontonian and threshold detection probabilities
    for N in (1, 2, 3, 4, 5, 6, 8, 10):
        cov = random_covariance(N, hbar=2)
        O = np.eye(2 * N) - np.linalg.inv(Qmat(cov, hbar=2))
        al = 0.3 * (rng.standard_normal(N) + 1j * rng.standard_normal(N))
        gamma = np.concatenate([al, al.conj()])
        out["ltor"].append({"N": N, "O": enc(O), "gamma": enc(gamma), "rec": enc(ltor(O, gamma)),
                            "direct": enc(numba_ltor(O, gamma))})
    for M in (2, 3, 4, 6):
        cov = random_covariance(M, hbar=2)
        mu = 0.5 * rng.standard_normal(2 * M)
        for _ in range(4):
            det = [int(x) for x in rng.integers(0, 2, M)]
            out["threshold"].append({"M": M, "cov": cov.tolist(), "mu": mu.tolist(), "det": det,
                                     "displaced": float(threshold_detection_prob(mu, cov, np.array(det))),
                                     "zero_mean": float(threshold_detection_prob(0 * mu, cov, np.array(det)))})
    # Bristolian
    for (m, n) in ((1, 1), (2, 2), (3, 2), (3, 3), (2, 4), (3, 5), (4, 4), (4, 6), (5, 7), (6, 8), (8, 10), (5, 4)):
        A = (rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))) / np.sqrt(m)
        Eh = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        E = 0.1 * (Eh @ Eh.conj().T)
        out["brs"].append({"m": m, "n": n, "A": enc(A), "E": enc(E), "value": enc(brs(A, E))})
        out["ubrs"].append({"m": m, "n": n, "A": enc(A), "value": enc(ubrs(A))})
    for M in (3, 4, 5):
        U = random_interferometer(M)
        T = np.sqrt(0.8) * U
        for _ in range(3):
            nin = [int(x) for x in rng.integers(0, 3, M)]
            nin[0] = max(nin[0], 1)                      # the reference indexes with an empty float array otherwise
            d = [int(x) for x in rng.integers(0, 2, M)]
            d[int(rng.integers(0, M))] = 1
            out["fock_threshold"].append({"M": M, "U": enc(U), "n": nin, "d": d,
                                          "unitary": float(fock_threshold_prob(nin, d, U)),
                                          "lossy": float(fock_threshold_prob(nin, d, T))})
    with open(os.path.join(HERE, "reference_outputs_next.json"), "w") as fh:
        json.dump(out, fh)
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
