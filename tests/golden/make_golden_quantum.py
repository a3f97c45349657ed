"""Reference outputs for the Fock-space callers of the hot path (thewalrus/quantum/fock_tensors.py:45-300):
pure_state_amplitude, state_vector, density_matrix (with and without post-selection), density_matrix_element.
Run once in the authoring container; tests/golden/reference_quantum.json is committed."""
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

from thewalrus.quantum import (density_matrix, density_matrix_element, pure_state_amplitude,  # noqa: E402
                               state_vector)
from thewalrus.random import random_covariance  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def enc(z):
    z = np.asarray(z, dtype=np.complex128)
    return {"re": z.real.tolist(), "im": z.imag.tolist()}


def main():
    out = {"pure": [], "mixed": []}
    rng = np.random.default_rng(20261020)
    for case, (M, displaced) in enumerate(((2, False), (2, True), (3, True))):
        np.random.seed(40 + case)
        cov = random_covariance(M, hbar=2, pure=True)
        mu = 0.5 * rng.standard_normal(2 * M) if displaced else np.zeros(2 * M)
        amps = []
        for _ in range(6):
            i = [int(v) for v in rng.integers(0, 4, M)]
            amps.append({"i": i, "with_pref": enc(pure_state_amplitude(mu, cov, i)),
                         "no_pref": enc(pure_state_amplitude(mu, cov, i, include_prefactor=False))})
        rec = {"M": M, "cov": cov.tolist(), "mu": mu.tolist(), "amps": amps,
               "state_vector": enc(state_vector(mu, cov, cutoff=4)),
               "state_vector_norm": enc(state_vector(mu, cov, cutoff=4, normalize=True))}
        ps = {M - 1: 1}
        rec["post_select"] = {str(k): v for k, v in ps.items()}
        rec["state_vector_ps"] = enc(state_vector(mu, cov, post_select=ps, cutoff=4))
        rec["state_vector_ps_norm"] = enc(state_vector(mu, cov, post_select=ps, cutoff=4, normalize=True))
        out["pure"].append(rec)
    for case, (M, displaced) in enumerate(((1, True), (2, False), (2, True))):
        np.random.seed(50 + case)
        cov = random_covariance(M, hbar=2, pure=False)
        mu = 0.4 * rng.standard_normal(2 * M) if displaced else np.zeros(2 * M)
        els = []
        for _ in range(6):
            i = [int(v) for v in rng.integers(0, 3, M)]
            j = [int(v) for v in rng.integers(0, 3, M)]
            els.append({"i": i, "j": j, "value": enc(density_matrix_element(mu, cov, i, j))})
        rec = {"M": M, "cov": cov.tolist(), "mu": mu.tolist(), "elements": els,
               "density_matrix": enc(density_matrix(mu, cov, cutoff=3))}
        if M > 1:
            ps = {0: 2}
            rec["post_select"] = {str(k): v for k, v in ps.items()}
            rec["density_matrix_ps"] = enc(density_matrix(mu, cov, post_select=ps, cutoff=3))
            rec["density_matrix_ps_norm"] = enc(density_matrix(mu, cov, post_select=ps, cutoff=3, normalize=True))
        out["mixed"].append(rec)
    with open(os.path.join(HERE, "reference_quantum.json"), "w") as fh:
        json.dump(out, fh)
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
