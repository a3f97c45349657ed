"""Generate golden vectors from the REFERENCE ITSELF (XanaduAI/thewalrus at /root/reference).

Run once in the authoring container (`python tests/golden/make_golden.py`); the GPU box has no
/root/reference, so the outputs are committed as tests/golden/*.json.  `dask` (the reference's only
missing dependency, unused on this path) is stubbed in sys.modules.
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
from thewalrus import charpoly  # noqa: E402
from thewalrus._hafnian import _haf, loop_hafnian, recursive_hafnian  # noqa: E402
from thewalrus._torontonian import numba_tor, rec_torontonian  # noqa: E402
from thewalrus.loop_hafnian_batch import loop_hafnian_batch  # noqa: E402
from thewalrus.random import random_covariance, random_interferometer  # noqa: E402
from thewalrus.quantum import Qmat, Amat, density_matrix_element, probabilities  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def enc(z):
    z = np.asarray(z)
    if np.iscomplexobj(z):
        return {"re": z.real.tolist(), "im": z.imag.tolist()}
    return {"re": z.astype(float).tolist(), "im": (0 * z.astype(float)).tolist()}


def sym(rng, n, kind):
    if kind == "complex":
        G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    elif kind == "real":
        G = rng.standard_normal((n, n))
    else:
        G = rng.integers(0, 2, (n, n)).astype(np.int64)
    return G + G.T


def main():
    out = {"hafnian": [], "loop_hafnian": [], "hafnian_repeated": [], "loop_hafnian_reps": [], "perm": [],
           "tor": [], "powertrace": [], "batch": [], "int_hafnian": [], "dme": []}
    rng = np.random.default_rng(20261017)
    # hafnian / loop hafnian, all reps 1
    for n in (2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22):
        for kind in ("complex", "real"):
            A = sym(rng, n, kind) / np.sqrt(n)
            out["hafnian"].append({"n": n, "kind": kind, "A": enc(A), "glynn": enc(_haf(A, glynn=True)),
                                   "inclexcl": enc(_haf(A, glynn=False))})
            out["loop_hafnian"].append({"n": n, "kind": kind, "A": enc(A), "value": enc(loop_hafnian(A))})
    for n in (5, 7, 9, 11):
        A = sym(rng, n, "complex") / np.sqrt(n)
        out["loop_hafnian"].append({"n": n, "kind": "complex", "A": enc(A), "value": enc(loop_hafnian(A))})
    # exact integer perfect-matching counts
    for n in (8, 12, 16, 20, 24, 28):
        G = (rng.random((n, n)) < 0.6).astype(np.int64)
        A = np.triu(G, 1)
        A = A + A.T
        out["int_hafnian"].append({"n": n, "A": A.tolist(), "value": int(recursive_hafnian(A))})
    # repeated rows
    for trial in range(24):
        n = int(rng.integers(2, 7))
        A = sym(rng, n, "complex") / np.sqrt(n)
        mu = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        rpt = [int(r) for r in rng.integers(0, 4, n)]
        if sum(rpt) == 0:
            rpt[0] = 2
        for glynn in (True, False):
            out["loop_hafnian_reps"].append({"A": enc(A), "mu": enc(mu), "rpt": rpt, "glynn": glynn,
                                             "value": enc(loop_hafnian(A, mu, rpt, glynn=glynn))})
            if sum(rpt) % 2 == 0:
                out["hafnian_repeated"].append({"A": enc(A), "rpt": rpt, "glynn": glynn,
                                                "value": enc(_haf(A, rpt, glynn=glynn))})
    # permanents
    for n in (4, 5, 8, 11, 14, 16):
        U = random_interferometer(2 * n)[:n, :n] if n > 4 else random_interferometer(n)
        R = rng.standard_normal((n, n))
        Z = rng.integers(-2, 4, (n, n)).astype(np.int64)
        out["perm"].append({"n": n, "kind": "complex", "A": enc(U), "bbfg": enc(thewalrus.perm(U, "bbfg")),
                            "ryser": enc(thewalrus.perm(U, "ryser"))})
        out["perm"].append({"n": n, "kind": "real", "A": enc(R), "bbfg": enc(thewalrus.perm(R, "bbfg")),
                            "ryser": enc(thewalrus.perm(R, "ryser"))})
        out["perm"].append({"n": n, "kind": "int", "A": Z.tolist(), "bbfg": float(thewalrus.perm(Z, "bbfg")),
                            "ryser": int(thewalrus.perm(Z, "ryser"))})
    # torontonians
    for N in (1, 2, 3, 4, 5, 6, 8, 10, 12):
        cov = random_covariance(N, hbar=2)
        O = np.eye(2 * N) - np.linalg.inv(Qmat(cov, hbar=2))
        out["tor"].append({"N": N, "kind": "complex", "O": enc(O), "rec": enc(rec_torontonian(O)),
                           "direct": enc(numba_tor(O)) if N <= 10 else None})
        B = rng.standard_normal((2 * N, 2 * N))
        Or = 0.95 * (B @ B.T) / np.linalg.norm(B @ B.T, 2)
        out["tor"].append({"N": N, "kind": "real", "O": enc(Or), "rec": enc(rec_torontonian(Or)),
                           "direct": enc(numba_tor(Or)) if N <= 10 else None})
    # power traces incl. the La Budde extension
    for (s, K) in ((2, 6), (4, 4), (4, 9), (6, 5), (6, 12), (8, 20)):
        H = rng.standard_normal((s, s)) + 1j * rng.standard_normal((s, s))
        out["powertrace"].append({"H": enc(H), "K": K, "value": enc(charpoly.powertrace(H.copy(), K))})
    # batched loop hafnian
    for trial in range(16):
        n = int(rng.integers(2, 6))
        A = sym(rng, n, "complex") / np.sqrt(n)
        D = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        fixed = [int(r) for r in rng.integers(0, 3, n - 1)]
        cutoff = int(rng.integers(1, 6))
        for glynn in (True, False):
            out["batch"].append({"A": enc(A), "D": enc(D), "fixed": fixed, "cutoff": cutoff, "glynn": glynn,
                                 "value": enc(loop_hafnian_batch(A, D, fixed, cutoff, glynn=glynn))})
    # GBS probabilities (caller of the batched front end): 4-mode displaced Gaussian state
    M = 4
    cov = random_covariance(M, hbar=2)
    mu = 0.3 * rng.standard_normal(2 * M)
    pats = [[int(x) for x in rng.integers(0, 3, M)] for _ in range(24)]
    out["dme"] = {"M": M, "cov": cov.tolist(), "mu": mu.tolist(), "patterns": pats,
                  "displaced": [float(density_matrix_element(mu, cov, p, p).real) for p in pats],
                  "zero_mean": [float(density_matrix_element(0 * mu, cov, p, p).real) for p in pats]}
    # host-side state preparation of the same state (pins thewalrus_b200.quantum.Qmat / Amat / prefactor)
    from thewalrus.quantum.conversions import complex_to_real_displacements  # noqa: E402
    from thewalrus.quantum.fock_tensors import _prefactor  # noqa: E402

    out["conv"] = {"Q": enc(Qmat(cov, hbar=2)), "A": enc(Amat(cov, hbar=2)),
                   "beta": enc(complex_to_real_displacements(mu, hbar=2)),
                   "prefactor": enc(_prefactor(mu, cov, hbar=2))}
    # BASELINE config 3 state (16 modes, bench.make_gbs_state): reference probabilities of a pattern sample
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import bench  # noqa: E402

    mu16, cov16, pats16 = bench.make_gbs_state(16, 2000, seed=3016)
    order = np.argsort(pats16.sum(axis=1), kind="stable")
    pick = [int(order[i]) for i in (0, 5, 50, 300, 700, 1000, 1300, 1600, 1800, 1900, 1950, 1990, 1995, 1999)]
    out["dme16"] = {"seed": 3016, "B": 2000, "index": pick, "patterns": [pats16[i].tolist() for i in pick],
                    "mu": mu16.tolist(), "cov": cov16.tolist(),
                    "displaced": [float(density_matrix_element(mu16, cov16, list(pats16[i]), list(pats16[i])).real)
                                  for i in pick],
                    "zero_mean": [float(density_matrix_element(0 * mu16, cov16, list(pats16[i]), list(pats16[i])).real)
                                  for i in pick]}
    with open(os.path.join(HERE, "reference_outputs.json"), "w") as fh:
        json.dump(out, fh)
    print("wrote", os.path.join(HERE, "reference_outputs.json"), {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
