"""The tensor-core (DMMA) batched loop-hafnian kernel (thewalrus_b200/csrc/pat_dmma.cuh) against the oracle and against
the warp-per-subset DFMA kernel it replaces, over every tile-shape class (E = 1 .. 16 matched edges), Glynn and
inclusion/exclusion, with and without loops, repeated vertices (delta = 0 deletions), tables of loop vectors and of
matrices.  Reference behaviour: loop_hafnian(A, D, reps) / hafnian_repeated, thewalrus/_hafnian.py:470-631."""
import os

import numpy as np
import pytest

import thewalrus_b200 as wb
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu


def _mat(rng, nv):
    G = rng.standard_normal((nv, nv)) + 1j * rng.standard_normal((nv, nv))
    return (G + G.T) / np.sqrt(2.0 * nv), (rng.standard_normal(nv) + 1j * rng.standard_normal(nv)) / np.sqrt(nv)


def _dfma(A, D, rpt, glynn=True, **kw):
    os.environ["WB200_PAT_DFMA"] = "1"
    try:
        return wb.quantum.lhaf_patterns(A, D, rpt, glynn, **kw)
    finally:
        del os.environ["WB200_PAT_DFMA"]


def _close(a, b, tol=1e-10):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-6 * np.max(np.abs(b)) + 1e-300)))


@pytest.mark.parametrize("E", list(range(1, 17)))
def test_every_tile_class_all_ones_edges(E):
    """2E distinct vertices with one repetition each: E edges, 2^(E-1) Glynn subsets — one pattern per class shape."""
    rng = np.random.default_rng(100 + E)
    nv = 2 * E + 3
    A, D = _mat(rng, nv)
    rpt = np.zeros((6, nv), dtype=np.int32)
    for b in range(6):
        rpt[b, rng.permutation(nv)[: 2 * E]] = 1
    for loops in (True, False):
        got = wb.quantum.lhaf_patterns(A, D if loops else None, rpt)
        ref = _dfma(A, D if loops else None, rpt)
        assert _close(got, ref) < 1e-10, (E, loops)
        if E <= 9:
            assert _close(got, co.lhaf_patterns(A, D if loops else None, rpt)) < 1e-10, (E, loops)


@pytest.mark.parametrize("glynn", [True, False])
@pytest.mark.parametrize("loops", [True, False])
def test_repeated_vertices_vs_oracle(glynn, loops):
    """Random repetition patterns (0..3 per vertex): integer delta incl. zeros, binomial weights, T > E."""
    rng = np.random.default_rng(7)
    nv = 8
    A, D = _mat(rng, nv)
    rpt = rng.integers(0, 4, (400, nv)).astype(np.int32)
    rpt = rpt[(rpt.sum(axis=1) % 2 == 0) & (rpt.sum(axis=1) <= 16)]
    got = wb.quantum.lhaf_patterns(A, D if loops else None, rpt, glynn)
    want = co.lhaf_patterns(A, D if loops else None, rpt, glynn)
    assert _close(got, want) < 1e-10
    assert _close(got, _dfma(A, D if loops else None, rpt, glynn)) < 1e-10


@pytest.mark.parametrize("glynn", [True, False])
def test_odd_totals_unpaired_vertex_vs_oracle(glynn):
    """Odd photon totals (f_loop_odd, thewalrus/_hafnian.py:246-285): the loop row of the tensor-core kernel carries the
    chain of the unpaired vertex's row next to the chain of D; every tile class with an odd vertex on top."""
    rng = np.random.default_rng(17)
    nv = 9
    A, D = _mat(rng, nv)
    rpt = rng.integers(0, 3, (600, nv)).astype(np.int32)
    rpt = rpt[(rpt.sum(axis=1) % 2 == 1) & (rpt.sum(axis=1) <= 15)]
    got = wb.quantum.lhaf_patterns(A, D, rpt, glynn)
    want = co.lhaf_patterns(A, D, rpt, glynn)
    assert _close(got, want) < 1e-10
    assert _close(got, _dfma(A, D, rpt, glynn)) < 1e-10
    for E in (3, 6, 8, 10, 12):      # 2E + 1 distinct vertices: E edges of one repetition and an unpaired vertex
        nv2 = 2 * E + 4
        A2, D2 = _mat(rng, nv2)
        r2 = np.zeros((4, nv2), dtype=np.int32)
        for b in range(4):
            r2[b, rng.permutation(nv2)[: 2 * E + 1]] = 1
        g2 = wb.quantum.lhaf_patterns(A2, D2, r2, glynn)
        assert _close(g2, _dfma(A2, D2, r2, glynn)) < 1e-10, E
        if E <= 8:
            assert _close(g2, co.lhaf_patterns(A2, D2, r2, glynn)) < 1e-10, E


def test_tor_v4_experiment_matches_default_kernel():
    """The opt-in warp-autonomous torontonian (WB200_TOR_V4=1, rank-2 DMMA Schur complements) against the default kernel."""
    import bench

    for w in ("tor24", "tor32"):
        _, _, O = bench.make_input(w)
        want = wb.tor(O)
        os.environ["WB200_TOR_V4"] = "1"
        try:
            got = wb.tor(O)
        finally:
            del os.environ["WB200_TOR_V4"]
        assert abs(got - want) <= 1e-11 * abs(want), w


def test_mixed_batch_even_odd_and_trivial_patterns():
    """Even patterns (DMMA classes), odd ones (DFMA fallback), N = 0 and N = 1 early exits in one call; ragged chunk tails."""
    rng = np.random.default_rng(11)
    nv = 12
    A, D = _mat(rng, nv)
    rpt = rng.integers(0, 3, (3000, nv)).astype(np.int32)
    rpt[rng.random(rpt.shape) < 0.55] = 0
    rpt[0] = 0
    rpt[1] = 0
    rpt[1, 5] = 1
    rpt = rpt[rpt.sum(axis=1) <= 14]
    got = wb.quantum.lhaf_patterns(A, D, rpt)
    want = co.lhaf_patterns(A, D, rpt)
    assert _close(got, want) < 1e-10


def test_tables_of_loop_vectors_and_matrices():
    rng = np.random.default_rng(13)
    nv, nA, nG, B = 10, 3, 5, 500
    As = np.stack([_mat(rng, nv)[0] for _ in range(nA)])
    Gs = np.stack([_mat(rng, nv)[1] for _ in range(nG)])
    rpt = (rng.random((B, nv)) < 0.6).astype(np.int32)
    rpt[rpt.sum(axis=1) % 2 == 1, 0] += 1
    ai = rng.integers(0, nA, B).astype(np.int32)
    gi = rng.integers(0, nG, B).astype(np.int32)
    got = wb.quantum.lhaf_patterns(As, Gs, rpt, gamma_index=gi, A_index=ai)
    want = np.array([co.lhaf_patterns(As[a], Gs[g], r[None, :])[0] for a, g, r in zip(ai, gi, rpt)])
    assert _close(got, want) < 1e-10


def test_gbs16_sample_matches_dfma_kernel():
    """BASELINE config 3 inputs: 4000 patterns of the 16-mode state, DMMA path vs the DFMA kernel of round 1."""
    import bench

    M, mu, cov, pats, A, gamma, rpt = bench.gbs_inputs("gbs16", 4000)
    got = wb.quantum.lhaf_patterns(A, gamma, rpt)
    ref = _dfma(A, gamma, rpt)
    assert _close(got, ref) < 1e-10
