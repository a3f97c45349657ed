"""Batched loop hafnian for many loop vectors — drop-in for
thewalrus.loop_hafnian_batch_gamma.loop_hafnian_batch_gamma (thewalrus/loop_hafnian_batch_gamma.py:222-269).

``out[k] == loop_hafnian_batch(A, D[k], fixed_reps, N_cutoff)`` for every row ``k`` of ``D`` (the threshold-detector
sampler passes one displaced diagonal per retained heterodyne outcome, thewalrus/samples.py:461).  One GPU sweep
(``wb200_lhaf_batch_gamma_host``) covers all rows: the reduced matrix and its power traces are built once per
subset and only the loop terms are redone per row, as the reference does (:107-111, :189-193).
"""
import numpy as np

from . import _engine
from ._prep import dd_sum, matched_reps
from .loop_hafnian_batch import add_batch_edges_even, add_batch_edges_odd

__all__ = ["loop_hafnian_batch_gamma"]


def loop_hafnian_batch_gamma(A, D, fixed_reps, N_cutoff, glynn=True, *, group=None, device=None):
    """Loop hafnians for every photon number 0..N_cutoff of the last mode and every row of ``D``.

    Same arguments, assertions and output (``complex128[n_D, N_cutoff + 1]``) as the reference.  ``group``/``device``
    as in :func:`thewalrus_b200.hafnian` (the subset index is sharded, partial tables all-reduced).
    """
    n = A.shape[0]
    assert A.shape[1] == n
    assert D.shape[1] == n
    assert len(fixed_reps) == n - 1
    N_cutoff = int(N_cutoff)
    n_D = D.shape[0]

    nz = np.nonzero(list(fixed_reps) + [1])[0]
    Anz = A[np.ix_(nz, nz)]
    Dnz = D[:, nz]
    fixed_nz = np.asarray(fixed_reps)[nz[:-1]]
    fixed_edges, fixed_m_reps, oddmode = matched_reps([int(r) for r in fixed_nz])

    if oddmode is None:
        batch_max, extra, odd_variant = N_cutoff // 2, N_cutoff % 2, 0
        edges = add_batch_edges_even(fixed_edges)
        edge_reps = np.concatenate(([batch_max], fixed_m_reps)).astype(np.int32)
        n_fixed = 2 * int(np.sum(fixed_m_reps))
        length = 2 * batch_max + extra + 1
    else:
        batch_max, extra, odd_variant = (N_cutoff - 1) // 2, 1 - (N_cutoff % 2), 1
        if batch_max < 0:
            # N_cutoff = 0 with an odd fixed photon number: the reference's sweep is empty and returns zeros
            # (SURVEY appendix); the only meaningful entry is the loop hafnian with the batch mode empty
            from ._hafnian import loop_hafnian

            return np.array([[loop_hafnian(A, D[k], list(fixed_reps) + [0], glynn=glynn, group=group, device=device)]
                             for k in range(n_D)], dtype=np.complex128)
        edges = add_batch_edges_odd(fixed_edges, oddmode)
        edge_reps = np.concatenate(([batch_max, 1], fixed_m_reps)).astype(np.int32)
        n_fixed = 2 * int(np.sum(fixed_m_reps)) + 1
        length = 2 * batch_max + extra + 2
    Ax = Anz[np.ix_(edges, edges)].astype(np.complex128)
    Dx = np.ascontiguousarray(Dnz[:, edges].astype(np.complex128))
    steps = int(np.prod(edge_reps.astype(object) + 1))

    def runner(lo, hi):
        return _engine.lhaf_batch_gamma_range(Ax, Dx, edge_reps, odd_variant, extra, glynn, lo, hi, length, device)

    table = _engine.run_sharded(steps, runner, group, width=4 * length * n_D)   # [world, 4 * n_D * length]
    table = table.reshape(table.shape[0], n_D, length, 4)
    out = np.empty((n_D, length), dtype=np.complex128)
    for k in range(n_D):
        for j in range(length):
            re, re_lo = dd_sum([(t[k, j, 0], t[k, j, 1]) for t in table])
            im, im_lo = dd_sum([(t[k, j, 2], t[k, j, 3]) for t in table])
            out[k, j] = complex(re + re_lo, im + im_lo)
            if glynn:
                out[k, j] *= 0.5 ** ((n_fixed + j) // 2)   # thewalrus/loop_hafnian_batch_gamma.py:119-122, 215-218
    return out
