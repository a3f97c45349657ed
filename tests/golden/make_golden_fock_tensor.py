"""Reference outputs for fock_tensor (Choi trick, thewalrus/quantum/fock_tensors.py:303-389) and the loss / noise
updates of photon-number distributions (:432-538).  Run once in the authoring container;
tests/golden/reference_fock_tensor.json is committed."""
import json, os, sys, types
import numpy as np
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")
_d = types.ModuleType("dask"); _d.delayed = lambda f, *a, **k: f; _d.compute = lambda *a, **k: a; sys.modules["dask"] = _d
sys.path.insert(0, "/root/reference")
from thewalrus.quantum import fock_tensor, loss_mat, update_probabilities_with_loss, update_probabilities_with_noise, probabilities
from thewalrus.random import random_symplectic, random_covariance
from thewalrus.symplectic import squeezing, beam_splitter, expand
def enc(z):
    z = np.asarray(z, dtype=np.complex128); return {"re": z.real.tolist(), "im": z.imag.tolist()}
out = {"fock_tensor": [], "loss": []}
rng = np.random.default_rng(20261022)
np.random.seed(7)
for l, disp, passive in ((1, True, False), (2, True, False), (2, False, True), (1, False, False)):
    if passive:
        S = beam_splitter(0.7, 0.3)
    else:
        S = random_symplectic(l)
    alpha = (0.3 * (rng.standard_normal(l) + 1j * rng.standard_normal(l))) if disp else np.zeros(l, dtype=complex)
    for sf in (False, True):
        out["fock_tensor"].append({"S": S.tolist(), "alpha": enc(alpha), "cutoff": 3, "sf_order": sf,
                                   "value": enc(fock_tensor(S, alpha, 3, sf_order=sf))})
cov = random_covariance(2, hbar=2, pure=False); mu = np.array([0.2, -0.1, 0.3, 0.1])
p = probabilities(mu, cov, 4)
noise = [np.array([0.8, 0.15, 0.05]), np.array([0.9, 0.1])]
out["loss"].append({"probs": p.tolist(), "etas": [0.7, 0.4], "lossy": update_probabilities_with_loss([0.7, 0.4], p).tolist(),
                    "noise": [n.tolist() for n in noise], "noisy": update_probabilities_with_noise(noise, p).tolist(),
                    "loss_mat": loss_mat(0.3, 5).tolist()})
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_fock_tensor.json"), "w"))
print(len(out["fock_tensor"]))
