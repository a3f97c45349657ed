"""CPU restatement of the bookkeeping of the symmetric-half hafnian kernel (thewalrus_b200/csrc/hafnian_sym.cu), checked
against plain linear algebra and against the oracle's hafnian.

The kernel never forms a whole product B_(k+1) = B_k S A': the row panel of a vertex pair in tile T (4 pairs = 8 rows per
tile) computes only the columns of tiles >= T (and of the one-pair tail), the rest is read by symmetry.  The power traces
the reference takes from charpoly.powertrace (thewalrus/charpoly.py:301-327, used at thewalrus/_hafnian.py:205, 433-467)
are assembled from the COMPUTED entries only:
    tr(M^(k+1))  = sum_v delta_v W[v][sigma(v)]                                   (one entry per row, in the diagonal tile)
    tr(M^(2k+1)) = sum_v delta_v sum_(computed c) w(v, c) W[sigma(v)][c] Y_old[v][c]
    tr(M^(2k+2)) = sum_v delta_v sum_(computed c) w(v, c) W[sigma(v)][c] Y_new[v][c]
with W = B_(k+1), Y_old = B_k S, Y_new = B_(k+1) S and the weight w = 2 on strictly-upper tiles (their mirror images are
never computed) and 1 on diagonal tiles.  This file emulates exactly that in NumPy — including a size that runs
zero-padded in the next whole-tile shape — so the identity the CUDA kernel relies on is pinned without a GPU."""
import numpy as np
import pytest

from oracle import walrus_oracle as wo


def _matched(A):
    n = A.shape[0]
    x = list(range(n - 1, -1, -2)) + list(range(n - 2, -1, -2))
    return A[np.ix_(x, x)]


def _computed_mask(m, TF, tail):
    """w(v, c) for the padded shape of TF full tiles (+ a tail pair): vertex v = pair + half * M."""
    M = 4 * TF + tail
    n = 2 * M
    pair = np.arange(n) % M
    tile = np.where(pair < 4 * TF, pair // 4, TF)            # TF = the tail "tile"
    w = np.zeros((n, n))
    for v in range(n):
        for c in range(n):
            if tile[v] < TF:                                 # regular panel: tiles >= its own, and the tail columns
                if tile[c] == tile[v]:
                    w[v, c] = 1.0
                elif tile[c] > tile[v]:
                    w[v, c] = 2.0
            elif tile[c] == TF:                              # tail panel: only its own 2 x 2 block
                w[v, c] = 1.0
    return w, M


def _kernel_style_traces(Ap, m, TF, tail, delta):
    """Power traces tr(M^1 .. M^m) of M = A' S the way haf_sym_kernel assembles them; Ap is the TRUE matrix (m pairs),
    embedded in the shape's M pairs with zero rows / columns."""
    w, M = _computed_mask(m, TF, tail)
    n = 2 * M
    P = np.zeros((n, n), dtype=complex)
    idx = np.concatenate([np.arange(m), M + np.arange(m)])     # true vertex (i, half) -> padded position i + half * M
    P[np.ix_(idx, idx)] = Ap
    d = np.ones(M)
    d[:m] = delta
    dv = np.concatenate([d, d])                              # delta of a vertex = delta of its pair
    sigma = (np.arange(n) + M) % n
    computed = w > 0

    def times_S(B):                                          # (B S)[v][c] = delta_c B[v][sigma(c)]
        return B[:, sigma] * dv[None, :]

    nprod = (m - 1) // 2
    K = nprod + 1
    tr = np.zeros(m + 1, dtype=complex)
    tr[1] = sum(dv[v] * P[v, sigma[v]] for v in range(n))
    B = P.copy()
    for k in range(1, nprod + 1):
        Yold = times_S(B)
        W = Yold @ P                                         # the kernel computes only W[computed]; the rest by symmetry:
        Wc = np.where(computed, W, 0.0)
        Wsym = Wc + np.where(computed.T & ~computed, Wc.T, 0.0)
        assert np.allclose(Wsym, W, rtol=0, atol=1e-9 * np.abs(W).max()), "computed tiles + their mirror images = the product"
        Ynew = times_S(Wsym)
        tr[k + 1] = sum(dv[v] * W[v, sigma[v]] for v in range(n) if computed[v, sigma[v]])
        if 2 * k + 1 > K and 2 * k + 1 <= m:
            tr[2 * k + 1] = sum(dv[v] * np.sum(w[v] * W[sigma[v]] * Yold[v]) for v in range(n))
        if 2 * k + 2 > K and 2 * k + 2 <= m:
            tr[2 * k + 2] = sum(dv[v] * np.sum(w[v] * W[sigma[v]] * Ynew[v]) for v in range(n))
        B = Wsym
    return tr[1:]


def _series(tr, m):
    """[x^m] exp(sum_j tr_j x^j / (2 j))  (thewalrus/_hafnian.py:183-214, f), push form as on the device."""
    c = np.zeros(m + 1, dtype=complex)
    c[0] = 1.0
    for s in range(1, m + 1):
        c[s] = sum(0.5 * tr[i - 1] * c[s - i] for i in range(1, s + 1)) / s
    return c[m]


@pytest.mark.parametrize("n,TF,tail", [(16, 2, 0), (18, 2, 1), (14, 2, 0), (12, 2, 0), (10, 1, 1), (26, 3, 1)])
def test_traces_from_computed_tiles_equal_power_traces(n, TF, tail):
    rng = np.random.default_rng(100 + n)
    G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    Ap = _matched(G + G.T)
    m = n // 2
    X = np.block([[np.zeros((m, m)), np.identity(m)], [np.identity(m), np.zeros((m, m))]])
    for _ in range(3):
        delta = rng.choice([-1.0, 1.0], size=m)
        delta[0] = -1.0
        Mj = Ap @ (X @ np.diag(np.concatenate([delta, delta])))
        want = np.array([np.trace(np.linalg.matrix_power(Mj, j)) for j in range(1, m + 1)])
        got = _kernel_style_traces(Ap, m, TF, tail, delta)
        scale = np.abs(want).max()
        assert np.max(np.abs(got - want)) <= 1e-10 * scale, (n, np.max(np.abs(got - want)) / scale)


@pytest.mark.parametrize("n,TF,tail", [(10, 1, 1), (12, 2, 0), (14, 2, 0)])
def test_hafnian_from_kernel_style_traces_equals_oracle(n, TF, tail):
    """The complete Glynn sum (thewalrus/_hafnian.py:416-467) with the kernel's traces and series against the oracle."""
    rng = np.random.default_rng(200 + n)
    G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    A = G + G.T
    Ap = _matched(A)
    m = n // 2
    total = 0.0
    for j in range(1 << (m - 1)):
        delta = np.array([1.0 if (j >> (m - 1 - p)) & 1 else -1.0 for p in range(m)])
        tr = _kernel_style_traces(Ap, m, TF, tail, delta)
        sign = -1.0 if (m - bin(j).count("1")) & 1 else 1.0
        total += sign * _series(tr, m)
    got = total / (1 << (m - 1))
    want = wo.haf(A)
    assert abs(got - want) <= 1e-10 * abs(want)
