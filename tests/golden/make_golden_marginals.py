"""Reference outputs for tvd_cutoff_bounds, n_body_marginals, find_classical_subsystem, real_to_complex_displacements
(thewalrus/quantum/fock_tensors.py:541-668, conversions.py:172-190).  Run once in the authoring container;
tests/golden/reference_marginals.json is committed."""
import json, os, sys, types
import numpy as np
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")
_d = types.ModuleType("dask"); _d.delayed = lambda f, *a, **k: f; _d.compute = lambda *a, **k: a; sys.modules["dask"] = _d
sys.path.insert(0, "/root/reference")
from thewalrus.quantum import tvd_cutoff_bounds, n_body_marginals, find_classical_subsystem, real_to_complex_displacements, complex_to_real_displacements
from thewalrus.random import random_covariance
out = []
rng = np.random.default_rng(20261023)
for case, (M, displaced) in enumerate(((2, True), (3, False), (3, True))):
    np.random.seed(80 + case)
    cov = random_covariance(M, hbar=2, pure=(case == 1))
    mu = 0.4 * rng.standard_normal(2 * M) if displaced else np.zeros(2 * M)
    marg = n_body_marginals(mu, cov, 3, 2)
    beta = complex_to_real_displacements(mu)
    out.append({"M": M, "cov": cov.tolist(), "mu": mu.tolist(), "tvd": tvd_cutoff_bounds(mu, cov, 5).tolist(),
                "marginals": [m.tolist() for m in marg], "classical": int(find_classical_subsystem(cov)),
                "r2c": real_to_complex_displacements(beta).tolist()})
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_marginals.json"), "w"))
print(len(out))
