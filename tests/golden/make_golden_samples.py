"""Seeded sample streams of the REFERENCE samplers (/root/reference/thewalrus/samples.py), for the drop-in
test of thewalrus_b200.samples with batch=1 (same numpy.random consumption order).  Run once in the authoring
container (`python tests/golden/make_golden_samples.py`); tests/golden/reference_samples.json is committed.
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

from thewalrus import samples as rs  # noqa: E402
from thewalrus.quantum import probabilities  # noqa: E402
from thewalrus.random import random_covariance  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    out = {"hafnian": [], "torontonian": [], "graph": [], "decompose": []}
    rng = np.random.default_rng(20261019)
    case = 0
    for M, pure, displaced in ((2, True, False), (2, False, True), (3, False, False), (3, True, True), (4, False, True)):
        np.random.seed(100 + case)
        cov = random_covariance(M, hbar=2, pure=pure)
        # keep the photon numbers small so the chains stay cheap for the CPU oracle used in the non-GPU test
        cov = 0.35 * cov + 0.65 * np.identity(2 * M)
        mu = 0.4 * rng.standard_normal(2 * M) if displaced else None
        T, sqrtW = rs.decompose_cov(cov)
        out["decompose"].append({"cov": cov.tolist(), "T": T.tolist(), "sqrtW": sqrtW.tolist()})
        for cutoff, seed in ((3, 11), (4, 12)):
            np.random.seed(seed + case)
            s = rs.hafnian_sample_state(cov, 6, mean=mu, cutoff=cutoff, max_photons=12)
            out["hafnian"].append({"cov": cov.tolist(), "mu": None if mu is None else mu.tolist(), "cutoff": cutoff,
                                   "max_photons": 12, "seed": seed + case, "samples": np.asarray(s).tolist()})
        for fanout, cutoff, seed in ((3, 1, 21), (4, 2, 22)):
            np.random.seed(seed + case)
            s = rs.torontonian_sample_state(cov, 6, mu=mu, fanout=fanout, cutoff=cutoff, max_photons=M)
            out["torontonian"].append({"cov": cov.tolist(), "mu": None if mu is None else mu.tolist(), "fanout": fanout,
                                       "cutoff": cutoff, "max_photons": M, "seed": seed + case,
                                       "samples": np.asarray(s).tolist()})
        case += 1
    A = np.array([[0, 1, 1, 0], [1, 0, 1, 1], [1, 1, 0, 1], [0, 1, 1, 0]], dtype=float)
    np.random.seed(31)
    out["graph"].append({"A": A.tolist(), "n_mean": 0.8, "cutoff": 3, "seed": 31, "kind": "hafnian",
                         "samples": np.asarray(rs.hafnian_sample_graph(A, 0.8, samples=5, cutoff=3)).tolist()})
    np.random.seed(32)
    out["graph"].append({"A": A.tolist(), "n_mean": 0.8, "fanout": 3, "seed": 32, "kind": "torontonian",
                         "samples": np.asarray(rs.torontonian_sample_graph(A, 0.8, samples=5, fanout=3)).tolist()})
    # exact photon-number distribution of a small displaced mixed state, for the statistical test of the batched path
    np.random.seed(77)
    cov = 0.4 * random_covariance(2, hbar=2, pure=False) + 0.6 * np.identity(4)
    mu = np.array([0.3, -0.2, 0.1, 0.25])
    out["distribution"] = {"cov": cov.tolist(), "mu": mu.tolist(), "cutoff": 6,
                           "probs": probabilities(mu, cov, 6).tolist()}
    with open(os.path.join(HERE, "reference_samples.json"), "w") as fh:
        json.dump(out, fh)
    print({k: (len(v) if isinstance(v, list) else "dict") for k, v in out.items()})


if __name__ == "__main__":
    main()
