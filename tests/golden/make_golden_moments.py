"""Reference outputs for the photon-number / click statistics callers (thewalrus/quantum/means_and_variances.py).
Run once in the authoring container; tests/golden/reference_moments.json is committed."""
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

from thewalrus import quantum as rq  # noqa: E402
from thewalrus.random import random_covariance  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def c(z):
    z = complex(z)
    return [z.real, z.imag]


def main():
    out = []
    rng = np.random.default_rng(20261021)
    for case, (M, displaced) in enumerate(((2, False), (3, True), (4, True))):
        np.random.seed(60 + case)
        cov = random_covariance(M, hbar=2, pure=False)
        mu = 0.5 * rng.standard_normal(2 * M) if displaced else np.zeros(2 * M)
        rec = {"M": M, "cov": cov.tolist(), "mu": mu.tolist(),
               "covmat": rq.photon_number_covmat(mu, cov).tolist(),
               "mean_clicks": float(rq.mean_clicks(cov)), "variance_clicks": float(rq.variance_clicks(cov)),
               "s_ordered": [], "expectation": [], "squared": [], "moment": [], "cumulant": [], "click_cumulant": []}
        for _ in range(5):
            rpt = [int(v) for v in rng.integers(0, 3, 2 * M)]
            s = float(rng.choice([-1.0, 0.0, 0.5, 1.0]))
            rec["s_ordered"].append({"rpt": rpt, "s": s, "value": c(rq.s_ordered_expectation(mu, cov, rpt, s=s))})
        for modes in ([0], [0, 1], list(range(M))):
            rec["expectation"].append({"modes": modes, "value": c(rq.photon_number_expectation(mu, cov, modes))})
            rec["squared"].append({"modes": modes, "value": c(rq.photon_number_squared_expectation(mu, cov, modes))})
        for ind in ({0: 2}, {0: 1, 1: 3}, {M - 1: 2, 0: 2}):
            rec["moment"].append({"indices": {str(k): v for k, v in ind.items()},
                                  "value": c(rq.photon_number_moment(mu, cov, ind))})
        for modes in ([0, 1], [0, 0, 1], [0, 1, M - 1, 1]):
            rec["cumulant"].append({"modes": modes, "value": c(rq.photon_number_cumulant(mu, cov, modes))})
        for modes in ([0, 1], list(range(M))):
            rec["click_cumulant"].append({"modes": modes, "value": c(rq.click_cumulant(mu, cov, modes))})
        out.append(rec)
    with open(os.path.join(HERE, "reference_moments.json"), "w") as fh:
        json.dump(out, fh)
    print(len(out), "states")


if __name__ == "__main__":
    main()
