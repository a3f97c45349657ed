"""Chain-rule GBS samplers on the batched loop-hafnian kernel.

Drop-in for the photon-number and threshold samplers of ``thewalrus.samples`` (thewalrus/samples.py:204-369,
407-620, 625-760).  The algorithm is the one of Bulmer et al. (arXiv:2108.01622) the reference implements: split
the state into a pure part and classical noise, draw a heterodyne outcome for all modes, then walk through the
modes undoing the heterodyne measurement one mode at a time and drawing that mode's photon number from loop
hafnians of the leading block of ``B``.

What is different here is the shape of the work.  The reference runs one chain at a time and, per mode, one
``loop_hafnian_batch`` call.  Here ``S`` chains advance together: one mode step is ONE call of the batched
front end (``wb200_lhaf_patterns_multi_host``) that evaluates ``loop_hafnian(B[:m, :m], gamma_s[:m],
reps = n_s[:m-1] + [k])`` for every chain ``s`` (every fan-out channel for the threshold sampler) and every
outcome ``k = 0..cutoff`` — the numbers ``loop_hafnian_batch(...)[k]`` / ``loop_hafnian_batch_gamma(...)[c, k]``
the reference draws from.  ``batch=1`` walks one chain at a time and consumes ``numpy.random``
exactly in the reference's order, so a seeded run reproduces the reference's samples; the default batches all
requested samples (same distribution, different stream).
"""
import numpy as np
from scipy.special import gammaln

from . import _engine
from . import quantum as _q

__all__ = ["decompose_cov", "mu_to_alpha", "invert_permutation", "photon_means_order", "get_heterodyne_fanout",
           "generate_hafnian_sample", "hafnian_sample_state", "hafnian_sample_graph",
           "generate_torontonian_sample", "torontonian_sample_state", "torontonian_sample_graph",
           "hafnian_sample_classical_state", "torontonian_sample_classical_state", "photon_number_sampler",
           "seed", "hafnian_sample_graph_rank_one"]


def seed(seed_val=None):
    """Seed ``numpy.random``, the generator every sampler here draws from (samples.py:720-729)."""
    np.random.seed(seed_val)


# ---------------------------------------------------------------------------------------------------
# small host helpers (samples.py:95-201)
# ---------------------------------------------------------------------------------------------------
def decompose_cov(cov, hbar=2):
    """cov = T + W with T = (hbar/2) S S^T pure and W = S (D - hbar/2) S^T classical noise; returns
    ``(T, sqrtW)`` with ``sqrtW sqrtW^T = W`` (samples.py:95-117)."""
    D, S = _q.williamson(cov)
    excess = np.diag(D) - hbar / 2
    excess = np.where(np.abs(excess) < 1e-10, 0.0, excess)
    return hbar / 2 * S @ S.T, S * np.sqrt(excess)


def mu_to_alpha(mu, hbar=2):
    """Complex amplitudes (x + i p) / sqrt(2 hbar) of xp means; works on a batch ``[..., 2M]`` too."""
    mu = np.asarray(mu)
    M = mu.shape[-1] // 2
    return (mu[..., :M] + 1j * mu[..., M:]) / np.sqrt(2 * hbar)


def invert_permutation(p):
    p = np.asarray(p)
    inv = np.empty(p.size, dtype=int)
    inv[p] = np.arange(p.size)
    return inv


def photon_means_order(mu, cov):
    """Modes sorted by increasing mean photon number, ties by index (samples.py:157-175)."""
    means = _q.photon_number_mean_vector(mu, cov)
    return np.asarray(sorted(range(len(means)), key=lambda i: (means[i], i)))


def _fanout(alpha, fanout):
    """``alpha[S, M]`` -> ``[S, M, fanout]``: every mode's amplitude in channel 0 of ``fanout`` channels, complex
    unit-variance noise in the others, mixed by a unitary DFT (samples.py:178-201).  Draw order per mode: the
    real parts, then the imaginary parts, as the reference."""
    S, M = alpha.shape
    chans = np.zeros((S, M, fanout), dtype=np.complex128)
    chans[:, :, 0] = alpha
    for j in range(M):
        re = np.random.normal(size=(S, fanout - 1))
        im = np.random.normal(size=(S, fanout - 1))
        chans[:, j, 1:] = re + 1j * im
    return np.fft.fft(chans, axis=2, norm="ortho")


def get_heterodyne_fanout(alpha, fanout):
    return _fanout(np.asarray(alpha, dtype=np.complex128)[None, :], fanout)[0]


# ---------------------------------------------------------------------------------------------------
# the batched chain
# ---------------------------------------------------------------------------------------------------
class _Chain:
    """State preparation shared by all chains of one Gaussian state (samples.py:228-247, 429-452)."""

    def __init__(self, cov, mu, hbar, scale=1.0):
        cov = np.asarray(cov)
        self.M = M = cov.shape[0] // 2
        mu = np.zeros(2 * M) if mu is None else np.asarray(mu)
        self.order = photon_means_order(mu, cov)
        self.order_inv = invert_permutation(self.order)
        oo = np.concatenate([self.order, self.order + M])
        self.mu = mu[oo]
        T, self.sqrtW = decompose_cov(cov[np.ix_(oo, oo)], hbar=hbar)
        self.chol = np.linalg.cholesky(T + np.identity(2 * M))
        self.B = np.ascontiguousarray(_q.Amat(T)[:M, :M] / scale)
        self.hbar = hbar

    def draw(self, S):
        """Pure-state means and heterodyne outcomes of S chains -> (pure_alpha, het_alpha), each [S, M]."""
        M = self.M
        pure_mu = self.mu + np.random.normal(size=(S, 2 * M)) @ self.sqrtW.T
        het_mu = pure_mu + np.random.normal(size=(S, 2 * M)) @ self.chol.T
        return mu_to_alpha(pure_mu, self.hbar), mu_to_alpha(het_mu, self.hbar)


def _engine_available():
    """The device walk bypasses ``quantum.lhaf_patterns``; tests that substitute the oracle for that call (CPU twins)
    keep the host walk."""
    return getattr(_q.lhaf_patterns, "__module__", "") == _q.__name__


def _lhaf_table(B, gammas, gidx, fixed, cutoff, device):
    """|lhaf|^2 / k! for k = 0..cutoff and every row of ``gidx``/``fixed``:
    ``loop_hafnian(B, gammas[gidx[r]], reps = fixed[r] + [k])`` -> ``[R, cutoff + 1]`` in one GPU call."""
    R, mode = fixed.shape
    K = cutoff + 1
    rpt = np.empty((R, K, mode + 1), dtype=np.int32)
    rpt[:, :, :mode] = fixed[:, None, :]
    rpt[:, :, mode] = np.arange(K)[None, :]
    lh = _q.lhaf_patterns(B, gammas, rpt.reshape(R * K, mode + 1), gamma_index=np.repeat(gidx, K), device=device)
    lh = lh.reshape(R, K)
    return (lh * lh.conj()).real * np.exp(-gammaln(np.arange(K) + 1.0))[None, :]


def _draw_outcomes(probs):
    """One outcome per row of ``probs[R, K]`` (unnormalised) by inverse CDF with one uniform per row — the draw
    ``numpy.random.choice(K, p=probs / probs.sum())`` makes (one ``random_sample``, ``searchsorted`` right)."""
    tot = probs.sum(axis=1, keepdims=True)
    if not (np.isfinite(probs).all() and (probs >= 0).all() and (tot > 0).all()):
        # numpy.random.choice(p=...) raises in the reference (samples.py:245-249) instead of drawing from garbage
        raise ValueError("probabilities contain NaN, are negative or do not sum to a positive number")
    p = probs / tot
    cdf = np.cumsum(p, axis=1)
    cdf /= cdf[:, -1:]
    u = np.random.random_sample(len(probs))
    return np.minimum((cdf <= u[:, None]).sum(axis=1), probs.shape[1] - 1)


DEVICE_CHAIN_MIN = 16      # from this many chains on, all mode steps run on the device (one C call per batch)


def _hafnian_chains(ch, S, cutoff, device):
    """Photon-number patterns (in the sorted mode order) of S chains advanced together.

    Small batches (and ``batch=1``, which reproduces the reference's seeded stream) walk the modes on the host with one
    GPU call per mode; from ``DEVICE_CHAIN_MIN`` chains on the whole walk — heterodyne shift, the S (cutoff + 1) loop
    hafnians of every mode step, the inverse-CDF draws — is one call of ``wb200_hafnian_chains_host`` and consumes
    ``numpy.random`` in the same order (normals, then one uniform per chain and mode, mode-major)."""
    M, B = ch.M, ch.B
    pure, het = ch.draw(S)
    gamma = pure.conj() + (het - pure) @ B.T
    if S >= DEVICE_CHAIN_MIN and cutoff <= 63 and _engine_available():
        u = np.random.random_sample((M, S))
        return _engine.hafnian_chains(B, gamma, het, u, cutoff, device)
    det = np.zeros((S, M), dtype=np.int32)
    sidx = np.arange(S, dtype=np.int32)
    for mode in range(M):
        m = mode + 1
        gamma = gamma - het[:, mode, None] * B[None, :, mode]
        probs = _lhaf_table(np.ascontiguousarray(B[:m, :m]), np.ascontiguousarray(gamma[:, :m]), sidx,
                            det[:, :mode], cutoff, device)
        det[:, mode] = _draw_outcomes(probs)
    return det


def _validate_cov(cov):
    if not isinstance(cov, np.ndarray):
        raise TypeError("Covariance matrix must be a NumPy array.")
    if cov.shape[0] != cov.shape[1]:
        raise ValueError("Covariance matrix must be square.")
    if np.isnan(cov).any():
        raise ValueError("Covariance matrix must not contain NaNs.")


def generate_hafnian_sample(cov, mean=None, hbar=2, cutoff=12, max_photons=8, *, device=None):
    """One photon-number sample, or -1 if it is rejected (last mode at the cutoff, or more than ``max_photons``
    photons) — samples.py:204-261; consumes ``numpy.random`` in the reference's order."""
    ch = _Chain(cov, mean, hbar)
    det = _hafnian_chains(ch, 1, cutoff, device)[0][ch.order_inv]
    if det[-1] == cutoff or det.sum() > max_photons:
        return -1
    return [int(v) for v in det]


def hafnian_sample_state(cov, samples, mean=None, hbar=2, cutoff=5, max_photons=30, parallel=False, *, batch=None,
                         device=None):
    """``samples`` photon-number samples of the Gaussian state (mean, cov) -> int array ``[samples, M]``
    (samples.py:264-369).  ``batch`` chains advance together per GPU call (default: all still-missing samples,
    at most 4096); ``batch=1`` reproduces the reference's seeded stream.  ``parallel`` is accepted for signature
    compatibility — the batch is the parallelism."""
    del parallel
    _validate_cov(cov)
    if batch is not None and int(batch) < 1:
        raise ValueError("batch must be >= 1")
    ch = _Chain(cov, mean, hbar)
    out, have = [], 0
    while have < samples:
        S = min(samples - have, 4096) if batch is None else int(batch)
        det = _hafnian_chains(ch, S, cutoff, device)[:, ch.order_inv]
        keep = (det[:, -1] != cutoff) & (det.sum(axis=1) <= max_photons)
        out.append(det[keep][: samples - have])
        have += len(out[-1])
    return np.concatenate(out).astype(int) if out else np.zeros((0, ch.M), dtype=int)


def hafnian_sample_graph(A, n_mean, samples=1, cutoff=5, max_photons=30, parallel=False, *, batch=None, device=None):
    """Samples of the pure GBS state encoding the graph ``A`` at mean photon number ``n_mean``
    (samples.py:372-399)."""
    cov = _q.Covmat(_q.adj_to_qmat(A, n_mean), hbar=2)
    return hafnian_sample_state(cov, samples, mean=None, hbar=2, cutoff=cutoff, max_photons=max_photons,
                                parallel=parallel, batch=batch, device=device)


def _torontonian_chains(ch, S, fanout, cutoff, max_photons, device):
    """Click patterns (sorted mode order) of S chains; returns (clicks[S, M], alive[S])."""
    M, B = ch.M, ch.B
    pure, het = ch.draw(S)
    hetf = _fanout(het, fanout)                                   # [S, M, fanout]
    gamma = pure.conj() / np.sqrt(fanout) + (hetf.sum(axis=2) - np.sqrt(fanout) * pure) @ B.T
    det = np.zeros((S, M), dtype=np.int32)
    clicks = np.zeros((S, M), dtype=np.int8)
    alive = np.ones(S, dtype=bool)
    for mode in range(M):
        m = mode + 1
        live = np.flatnonzero(alive)
        if live.size == 0:
            break
        # gamma after removing channels 0..c of this mode, for every live chain: [L, fanout, M]
        shifts = np.cumsum(hetf[live, mode, :], axis=1)
        gf = gamma[live, None, :] - shifts[:, :, None] * B[None, None, :, mode]
        L = live.size
        probs = _lhaf_table(np.ascontiguousarray(B[:m, :m]), np.ascontiguousarray(gf[:, :, :m]).reshape(L * fanout, m),
                            np.arange(L * fanout, dtype=np.int32), np.repeat(det[live, :mode], fanout, axis=0),
                            cutoff, device).reshape(L, fanout, cutoff + 1)
        waiting = np.arange(L)                                    # chains that have not clicked in this mode yet
        gamma[live] = gf[:, fanout - 1, :]
        for c in range(fanout):
            if waiting.size == 0:
                break
            o = _draw_outcomes(probs[waiting, c, :])
            hit = waiting[o > 0]
            det[live[hit], mode] += o[o > 0]
            clicks[live[hit], mode] = 1
            gamma[live[hit]] = gf[hit, c, :]
            waiting = waiting[o == 0]
        alive[live] &= clicks[live].sum(axis=1) <= max_photons
    return clicks, alive


def generate_torontonian_sample(cov, mu=None, hbar=2, max_photons=30, fanout=10, cutoff=1, *, device=None):
    """One threshold (click) sample, or -1 if more than ``max_photons`` detectors click — samples.py:407-481;
    consumes ``numpy.random`` in the reference's order."""
    ch = _Chain(cov, mu, hbar, scale=fanout)
    clicks, alive = _torontonian_chains(ch, 1, fanout, cutoff, max_photons, device)
    if not alive[0]:
        return -1
    return [int(v) for v in clicks[0][ch.order_inv]]


def torontonian_sample_state(cov, samples, mu=None, hbar=2, max_photons=30, fanout=10, cutoff=1, parallel=False, *,
                             batch=None, device=None):
    """``samples`` threshold samples of the Gaussian state (mu, cov) -> int array ``[samples, M]``
    (samples.py:484-588).  Batching as in :func:`hafnian_sample_state`."""
    del parallel
    _validate_cov(cov)
    if batch is not None and int(batch) < 1:
        raise ValueError("batch must be >= 1")
    ch = _Chain(cov, mu, hbar, scale=fanout)
    out, have = [], 0
    while have < samples:
        S = min(samples - have, 1024) if batch is None else int(batch)
        clicks, alive = _torontonian_chains(ch, S, fanout, cutoff, max_photons, device)
        out.append(clicks[alive][:, ch.order_inv][: samples - have])
        have += len(out[-1])
    return np.concatenate(out).astype(int) if out else np.zeros((0, ch.M), dtype=int)


def torontonian_sample_graph(A, n_mean, samples=1, max_photons=30, fanout=10, cutoff=1, parallel=False, *, batch=None,
                             device=None):
    """Threshold samples of the pure GBS state encoding the graph ``A`` (samples.py:591-620)."""
    cov = _q.Covmat(_q.adj_to_qmat(A, n_mean), hbar=2)
    return torontonian_sample_state(cov, samples, hbar=2, max_photons=max_photons, fanout=fanout, cutoff=cutoff,
                                    parallel=parallel, batch=batch, device=device)


# ---------------------------------------------------------------------------------------------------
# samplers that need no hafnian at all (host only; samples.py:625-760)
# ---------------------------------------------------------------------------------------------------
def hafnian_sample_classical_state(cov, samples, mean=None, hbar=2, atol=1e-08, cutoff=None):
    """Positive-P states: draw coherent amplitudes from the P function, then Poisson photon numbers."""
    del cutoff
    if not _q.is_classical_cov(cov, hbar=hbar, atol=atol):
        raise ValueError("Not a classical covariance matrix")
    n = cov.shape[0]
    if mean is None:
        mean = np.zeros(n)
    elif mean.shape != (n,):
        raise ValueError("mean and cov do not have compatible shapes")
    R = np.random.multivariate_normal(mean, cov - 0.5 * hbar * np.identity(n), samples)
    alpha = (R[:, : n // 2] + 1j * R[:, n // 2:]) / np.sqrt(2 * hbar)
    return np.random.poisson(np.abs(alpha) ** 2)


def torontonian_sample_classical_state(cov, samples, mean=None, hbar=2, atol=1e-08):
    return np.where(hafnian_sample_classical_state(cov, samples, mean=mean, hbar=hbar, atol=atol) > 0, 1, 0)


def photon_number_sampler(probabilities, num_samples, out_of_bounds=False):
    """Samples from a photon-number probability tensor of shape ``[cutoff] * modes`` (samples.py:683-717):
    renormalised if ``out_of_bounds`` is False, else the missing mass maps to the ``out_of_bounds`` placeholder."""
    modes, cutoff = probabilities.ndim, probabilities.shape[0]
    flat = probabilities.flatten()
    total = flat.sum()
    shape = [cutoff] * modes
    if out_of_bounds is False:
        vals = np.arange(flat.size, dtype=int)
        return [np.unravel_index(np.random.choice(vals, p=flat / total), shape) for _ in range(num_samples)]
    vals = np.arange(flat.size + 1, dtype=int)
    p = np.append(flat, 1.0 - total)
    picks = [np.random.choice(vals, p=p) for _ in range(num_samples)]
    return [out_of_bounds if i == flat.size else np.unravel_index(i, shape) for i in picks]


def hafnian_sample_graph_rank_one(G, n_mean, samples=1):
    """Rank-one adjacency matrix A = G G^T: the total photon number is twice a negative binomial, and the photons
    land independently on detectors with probability |G_i|^2 / sum |G|^2 (samples.py:732-767)."""
    G = np.asarray(G)
    q = 1.0 - np.tanh(np.arcsinh(np.sqrt(n_mean))) ** 2
    p = np.abs(G) ** 2
    p = p / p.sum()
    out = np.zeros((samples, len(G)))
    for s in range(samples):
        total = 2 * np.random.negative_binomial(0.5, q, 1)[0]
        for _ in range(total):
            out[s, np.random.choice(len(G), p=p)] += 1
    return out
