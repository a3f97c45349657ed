"""Batched loop hafnian — drop-in for thewalrus.loop_hafnian_batch.loop_hafnian_batch
(thewalrus/loop_hafnian_batch.py:260-304).

Returns ``[lhaf(A, D, reps = fixed_reps + [k]) for k in 0..N_cutoff]`` from ONE sweep over the subsets of
(batch edge + fixed edges); the sweep (``_calc_loop_hafnian_batch_even/_odd``, :51-208) runs on the GPU
through ``wb200_lhaf_batch_host``; the edge bookkeeping (:211-257) is integer work on the host.
"""
import numpy as np

from . import _engine
from ._prep import dd_sum, matched_reps

__all__ = ["loop_hafnian_batch", "add_batch_edges_even", "add_batch_edges_odd"]


def add_batch_edges_even(fixed_edges):
    """Vertex order with the batch mode first, paired with itself (thewalrus/loop_hafnian_batch.py:211-231)."""
    fixed_edges = np.asarray(fixed_edges, dtype=int)
    ne = len(fixed_edges)
    if ne == 0:
        return np.array([0, 0], dtype=int)
    new = int(fixed_edges.max()) + 1
    half = ne // 2
    return np.concatenate(([new], fixed_edges[:half], [new], fixed_edges[half:])).astype(int)


def add_batch_edges_odd(fixed_edges, oddmode):
    """Same with an unpaired fixed vertex: second edge (oddmode, batch mode), one repetition
    (thewalrus/loop_hafnian_batch.py:234-257)."""
    fixed_edges = np.asarray(fixed_edges, dtype=int)
    ne = len(fixed_edges)
    if ne == 0:
        return np.array([1, oddmode, 1, 1], dtype=int)
    new = max(int(fixed_edges.max()), int(oddmode)) + 1
    half = ne // 2
    return np.concatenate(([new, oddmode], fixed_edges[:half], [new, new], fixed_edges[half:])).astype(int)


def loop_hafnian_batch(A, D, fixed_reps, N_cutoff, glynn=True, *, group=None, device=None):
    """Loop hafnians for every photon number 0..N_cutoff of the last mode, the others fixed.

    Same arguments, assertions and output (``complex128[N_cutoff + 1]``) as the reference.  ``group``/``device``
    as in :func:`thewalrus_b200.hafnian` (the subset index is sharded, partial vectors all-reduced).
    """
    n = A.shape[0]
    assert A.shape[1] == n
    assert D.shape == (n,)
    assert len(fixed_reps) == n - 1
    N_cutoff = int(N_cutoff)

    nz = np.nonzero(list(fixed_reps) + [1])[0]
    Anz = A[np.ix_(nz, nz)]
    Dnz = D[nz]
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
            # the reference itself mis-handles N_cutoff = 0 with an odd fixed photon number:
            # the only output is lhaf with the batch mode empty
            from ._hafnian import loop_hafnian

            return np.array([loop_hafnian(A, D, list(fixed_reps) + [0], glynn=glynn, group=group, device=device)],
                            dtype=np.complex128)
        edges = add_batch_edges_odd(fixed_edges, oddmode)
        edge_reps = np.concatenate(([batch_max, 1], fixed_m_reps)).astype(np.int32)
        n_fixed = 2 * int(np.sum(fixed_m_reps)) + 1
        length = 2 * batch_max + extra + 2
    Ax = Anz[np.ix_(edges, edges)].astype(np.complex128)
    Dx = Dnz[edges].astype(np.complex128)
    steps = int(np.prod(edge_reps.astype(object) + 1))

    def runner(lo, hi):
        return _engine.lhaf_batch_range(Ax, Dx, edge_reps, odd_variant, extra, glynn, lo, hi, length, device)

    table = _engine.run_sharded(steps, runner, group, width=4 * length)   # [world, 4 * length]
    table = table.reshape(table.shape[0], length, 4)
    out = np.empty(length, dtype=np.complex128)
    for j in range(length):
        re, re_lo = dd_sum([(t[j, 0], t[j, 1]) for t in table])
        im, im_lo = dd_sum([(t[j, 2], t[j, 3]) for t in table])
        out[j] = complex(re + re_lo, im + im_lo)
        if glynn:
            out[j] *= 0.5 ** ((n_fixed + j) // 2)   # thewalrus/loop_hafnian_batch.py:118-121, 203-206
    return out
