"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.

A NumPy restatement of the reference's algorithms for the exponential-sum hot path
(XanaduAI/thewalrus v0.23.0-dev), written from the formulas, each function citing the reference lines it
follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product (``thewalrus_b200``) never does.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function here against
(i) the known-answer tests of the reference's own suite (thewalrus/tests/test_hafnian.py:182-299,
test_permanent.py:56-152, test_torontonian.py:75-126, test_labudde.py:34-74, test_hafnian_repeated.py) and
(ii) outputs of the reference itself, generated in the authoring container by
``tests/golden/make_golden.py`` (imports /root/reference) and committed as ``tests/golden/*.json``.

Every driver takes a half-open index range so that shards / sampled ranges can be checked too.
"""
import math

import numpy as np

# --------------------------------------------------------------------------------------------------
# integer helpers
# --------------------------------------------------------------------------------------------------


def matched_reps(reps):
    """thewalrus/_hafnian.py:80-159 — greedy pairing of repeated vertices."""
    n = len(reps)
    if sum(reps) == 0:
        return np.array([], dtype=np.int64), np.array([], dtype=np.int64), None
    live = sorted(((int(r), i) for i, r in enumerate(reps) if r > 0), reverse=True)
    a_side, b_side, mult = [], [], []
    while len(live) > 1 or (len(live) == 1 and live[0][0] > 1):
        live.sort(reverse=True)
        if len(live) == 1 or live[0][0] > 2 * live[1][0]:
            r, v = live[0]
            a_side.append(v)
            b_side.append(v)
            mult.append(r // 2)
            live = live[1:] if r % 2 == 0 else [(1, v)] + live[1:]
        else:
            (r0, v0), (r1, v1) = live[0], live[1]
            a_side.append(v0)
            b_side.append(v1)
            mult.append(r1)
            live = ([(r0 - r1, v0)] if r0 > r1 else []) + live[2:]
    odd = live[0][1] if len(live) == 1 else None
    return np.array(a_side + b_side, dtype=np.int64), np.array(mult, dtype=np.int64), odd


def find_kept_edges(j, reps):
    """thewalrus/_hafnian.py:162-180 — mixed-radix digits, most significant first."""
    digits = []
    num = int(j)
    for base in [int(r) + 1 for r in reps][::-1]:
        digits.append(num % base)
        num //= base
    return np.array(digits[::-1], dtype=np.int64)


def num_steps(edge_reps, glynn=True, has_odd=False):
    """thewalrus/_hafnian.py:432-435 and :535-538."""
    edge_reps = [int(e) for e in edge_reps]
    if glynn and not has_odd:
        return ((edge_reps[0] + 2) // 2) * int(np.prod([e + 1 for e in edge_reps[1:]], dtype=object))
    return int(np.prod([e + 1 for e in edge_reps], dtype=object))


# --------------------------------------------------------------------------------------------------
# power traces (thewalrus/charpoly.py)
# --------------------------------------------------------------------------------------------------


def hessenberg(H):
    """Householder reduction to upper Hessenberg form, in place (thewalrus/charpoly.py:42-125)."""
    n = len(H)
    for k in range(1, n - 1):
        col = H[k:, k - 1].copy()
        sigma = np.linalg.norm(col)
        if col[0] != 0:
            sigma = sigma * col[0] / abs(col[0])
        v = col
        v[0] += sigma
        nv = np.linalg.norm(v) ** 2
        if nv == 0:
            continue
        H[k:, k - 1:] -= 2.0 * np.outer(v, v.conj() @ H[k:, k - 1:]) / nv
        H[:, k:] -= 2.0 * np.outer(H[:, k:] @ v, v.conj()) / nv
    return H


def labudde(H):
    """Characteristic-polynomial coefficients c_1..c_n of an upper-Hessenberg matrix by La Budde's
    recurrence (thewalrus/charpoly.py:195-282, arXiv:1104.3769): p_i = (x - a_i) p_{i-1}
    - sum_m h_{i-m,i} (prod beta) p_{i-m-1}."""
    n = len(H)
    c = np.zeros((n + 1, n + 1), dtype=H.dtype)  # c[i][j]: coefficient of x^(i-j) in p_i, c[i][0] = 1
    c[:, 0] = 1
    for i in range(1, n + 1):
        a = H[i - 1, i - 1]
        for j in range(1, i + 1):
            val = (c[i - 1, j] if j <= i - 1 else 0) - a * c[i - 1, j - 1]
            bprod = 1.0
            for m in range(1, j):
                bprod = bprod * H[i - m, i - m - 1]
                val -= H[i - m - 1, i - 1] * bprod * c[i - m - 1, j - m - 1]
            c[i, j] = val
    return c[n, 1:]


def powertrace(H, K):
    """[size, tr H, ..., tr H^(K-1)] — thewalrus/charpoly.py:301-327: matrix products while the power is
    below the size, then Newton's recurrence on the La Budde characteristic polynomial."""
    H = np.array(H, dtype=np.complex128)
    m = len(H)
    if m == 0:
        return np.zeros(K, dtype=np.complex128)
    lim = min(K, m)
    tr = [complex(m), np.trace(H)]
    P = H
    for _ in range(lim - 2):
        P = P @ H
        tr.append(np.trace(P))
    tr = tr[:max(lim, 2)]
    if K <= m:
        return np.array(tr[:K], dtype=np.complex128)
    coeffs = labudde(hessenberg(H.copy()))
    while len(tr) < K:
        tr.append(-sum(coeffs[k] * tr[-k - 1] for k in range(m)))
    return np.array(tr, dtype=np.complex128)


def exp_series(factors, order):
    """Coefficients 0..order of exp(sum_i factors[i] eta^i), built factor by factor exactly as the `comb`
    loops of f / f_loop / f_loop_odd do (thewalrus/_hafnian.py:196-209, 228-242, 264-285)."""
    comb = np.zeros(order + 1, dtype=np.complex128)
    comb[0] = 1
    for i in range(1, order + 1):
        fac = factors[i]
        new = comb.copy()
        pw = 1.0
        for j in range(1, order // i + 1):
            pw = pw * fac / j
            new[i * j:] += comb[: order + 1 - i * j] * pw
        comb = new
    return comb


# --------------------------------------------------------------------------------------------------
# per-subset pieces
# --------------------------------------------------------------------------------------------------


def submatrices(delta, A, D=None, oddV=None):
    """thewalrus/_hafnian.py:288-356 — drop rows/cols with delta = 0, swap halves, scale columns."""
    delta = np.asarray(delta)
    keep = np.nonzero(delta)[0]
    k = len(keep)
    rows = np.concatenate((keep, keep + len(delta)))
    d = delta[keep].astype(np.complex128)
    A_nz = A[np.ix_(rows, rows)]
    AX = np.empty_like(A_nz, dtype=np.complex128)
    AX[:, :k] = d * A_nz[:, k:]
    AX[:, k:] = d * A_nz[:, :k]
    XD = Dk = oddVX = None
    if D is not None:
        Dk = D[rows].astype(np.complex128)
        XD = np.concatenate((d * Dk[k:], d * Dk[:k]))
    if oddV is not None:
        ov = oddV[rows]
        oddVX = np.concatenate((d * ov[k:], d * ov[:k]))
    return AX, XD, Dk, oddVX


def f(AX, n):
    """thewalrus/_hafnian.py:183-209."""
    order = n // 2
    pt = powertrace(AX, order + 1)
    fac = [0] + [pt[i] / (2 * i) for i in range(1, order + 1)]
    return exp_series(fac, order)


def f_loop(AX, XD, D, n):
    """thewalrus/_hafnian.py:212-242."""
    order = n // 2
    pt = powertrace(AX, order + 1)
    fac = [0]
    xd = XD
    for i in range(1, order + 1):
        fac.append(pt[i] / (2 * i) + (xd @ D) / 2 if len(D) else pt[i] / (2 * i))
        xd = xd @ AX
    return exp_series(fac, order)


def f_loop_odd(AX, XD, D, n, oddloop, oddVX):
    """thewalrus/_hafnian.py:246-285."""
    pt = powertrace(AX, n // 2 + 1)
    fac = [0]
    xd, dd = XD, D
    for i in range(1, n + 1):
        if i == 1:
            fac.append(oddloop)
        elif i % 2 == 0:
            fac.append(pt[i // 2] / i + (xd @ dd) / 2)
        else:
            fac.append(oddVX @ dd)
            dd = AX @ dd
            # the reference advances D_S only; XD_S is fixed (so xd @ dd = XD M^(t-1) D)
    return exp_series(fac, n)


# --------------------------------------------------------------------------------------------------
# subset-sum drivers
# --------------------------------------------------------------------------------------------------


def _binom_prod(edge_reps, kept, start=0):
    out = 1.0
    for r, k in zip(edge_reps[start:], kept[start:]):
        out *= math.comb(int(r), int(k))
    return out


def calc_hafnian(A, edge_reps, glynn=True, j0=0, j1=None, scale=True):
    """thewalrus/_hafnian.py:416-467 over subset indices [j0, j1)."""
    edge_reps = np.asarray(edge_reps, dtype=np.int64)
    N = 2 * int(edge_reps.sum())
    steps = num_steps(edge_reps, glynn)
    j1 = steps if j1 is None else j1
    H = 0j
    for j in range(j0, j1):
        kept = find_kept_edges(j, edge_reps)
        esum = int(kept.sum())
        w = _binom_prod(edge_reps, kept)
        delta = 2 * kept - edge_reps if glynn else kept
        AX, _, _, _ = submatrices(delta, A)
        pre = (-1.0) ** (N // 2 - esum) * w
        if glynn and delta[0] == 0:
            pre *= 0.5
        H += pre * f(AX, N)[N // 2]
    if glynn and scale:
        H *= 0.5 ** (N // 2 - 1)
    return H


def calc_loop_hafnian(A, D, edge_reps, oddloop=None, oddV=None, glynn=True, j0=0, j1=None, scale=True):
    """thewalrus/_hafnian.py:512-577 over subset indices [j0, j1)."""
    edge_reps = np.asarray(edge_reps, dtype=np.int64)
    N = 2 * int(edge_reps.sum()) + (1 if oddloop is not None else 0)
    steps = num_steps(edge_reps, glynn, oddloop is not None)
    j1 = steps if j1 is None else j1
    H = 0j
    for j in range(j0, j1):
        kept = find_kept_edges(j, edge_reps)
        esum = int(kept.sum())
        w = _binom_prod(edge_reps, kept)
        delta = 2 * kept - edge_reps if glynn else kept
        AX, XD, Dk, oddVX = submatrices(delta, A, D, oddV)
        pre = (-1.0) ** (N // 2 - esum) * w
        if oddloop is not None:
            H += pre * f_loop_odd(AX, XD, Dk, N, oddloop, oddVX)[N]
        else:
            if glynn and delta[0] == 0:
                pre *= 0.5
            H += pre * f_loop(AX, XD, Dk, N)[N // 2]
    if glynn and scale:
        H *= 0.5 ** (N // 2 if oddloop is not None else N // 2 - 1)
    return H


def haf(A, reps=None, glynn=True):
    """thewalrus/_hafnian.py:470-508."""
    n = A.shape[0]
    reps = [1] * n if reps is None else list(reps)
    N = sum(reps)
    if N == 0:
        return 1.0
    if N % 2:
        return 0.0
    x, edge_reps, _ = matched_reps(reps)
    Ax = A[np.ix_(x, x)].astype(np.complex128)
    return calc_hafnian(Ax, edge_reps, glynn)


def loop_hafnian(A, D=None, reps=None, glynn=True):
    """thewalrus/_hafnian.py:581-631."""
    n = A.shape[0]
    reps = [1] * n if reps is None else list(reps)
    D = A.diagonal() if D is None else np.asarray(D)
    N = sum(reps)
    if N == 0:
        return 1.0
    if N == 1:
        return D[[i for i, r in enumerate(reps) if r == 1][0]]
    x, edge_reps, odd = matched_reps(reps)
    oddloop = np.complex128(D[odd]) if odd is not None else None
    oddV = A[odd, x].astype(np.complex128) if odd is not None else None
    Ax = A[np.ix_(x, x)].astype(np.complex128)
    Dx = D[x].astype(np.complex128)
    return calc_loop_hafnian(Ax, Dx, edge_reps, oddloop, oddV, glynn)


# --------------------------------------------------------------------------------------------------
# batched loop hafnian (thewalrus/loop_hafnian_batch.py)
# --------------------------------------------------------------------------------------------------


def _batch_core(A, D, fixed_edge_reps, batch_max, cutoff_extra, glynn, odd_variant):
    """thewalrus/loop_hafnian_batch.py:51-123 (even) and :127-208 (odd)."""
    fixed_edge_reps = np.asarray(fixed_edge_reps, dtype=np.int64)
    oddloop, oddV = D[0], A[0, :]
    n = A.shape[0]
    if odd_variant:
        oddloop0, oddV0 = D[1], A[1, :]
        N_fixed = 2 * int(fixed_edge_reps.sum()) + 1
        N_max = N_fixed + 2 * batch_max + cutoff_extra + 1
        edge_reps = np.concatenate(([batch_max, 1], fixed_edge_reps)).astype(np.int64)
        length = 2 * batch_max + cutoff_extra + 2
    else:
        N_fixed = 2 * int(fixed_edge_reps.sum())
        N_max = N_fixed + 2 * batch_max + cutoff_extra
        edge_reps = np.concatenate(([batch_max], fixed_edge_reps)).astype(np.int64)
        length = 2 * batch_max + cutoff_extra + 1
    steps = int(np.prod(edge_reps + 1))
    H = np.zeros(length, dtype=np.complex128)
    for j in range(steps):
        kept = find_kept_edges(j, edge_reps)
        esum = int(kept.sum())
        w = _binom_prod(edge_reps, kept, start=1)
        delta = 2 * kept - edge_reps if glynn else kept
        AX, XD, Dk, oddVX = submatrices(delta, A, D, oddV)
        if odd_variant and kept[0] == 0 and kept[1] == 0:
            _, _, _, oddVX0 = submatrices(delta, A, D, oddV0)
            pm = (-1) ** (N_fixed // 2 - esum)
            H[0] += w * pm * f_loop_odd(AX, XD, Dk, N_fixed, oddloop0, oddVX0)[N_fixed]
        fe = f_loop(AX, XD, Dk, N_max)
        fo = f_loop_odd(AX, XD, Dk, N_max, oddloop, oddVX)
        first = 2 * int(kept[0]) + (1 if odd_variant else 0)
        for N_det in range(first, length):
            N = N_fixed + N_det
            pm = (-1.0) ** (N // 2 - esum)
            half = (N_det - 1) // 2 if odd_variant else N_det // 2
            wt = math.comb(half, int(kept[0])) * w
            H[N_det] += wt * pm * (fe[N // 2] if N % 2 == 0 else fo[N])
    if glynn:
        for j in range(length):
            H[j] *= 0.5 ** ((N_fixed + j) // 2)
    return H


def _batch_edges_even(fixed_edges):
    """thewalrus/loop_hafnian_batch.py:211-231."""
    if len(fixed_edges) == 0:
        return np.array([0, 0], dtype=int)
    ne = len(fixed_edges)
    new = max(fixed_edges) + 1
    return np.array([new] + list(fixed_edges[: ne // 2]) + [new] + list(fixed_edges[ne // 2:]), dtype=int)


def _batch_edges_odd(fixed_edges, oddmode):
    """thewalrus/loop_hafnian_batch.py:234-257."""
    if len(fixed_edges) == 0:
        return np.array([1, oddmode, 1, 1], dtype=int)
    ne = len(fixed_edges)
    new = max(max(fixed_edges), oddmode) + 1
    return np.array([new, oddmode] + list(fixed_edges[: ne // 2]) + [new, new] + list(fixed_edges[ne // 2:]),
                    dtype=int)


def loop_hafnian_batch(A, D, fixed_reps, N_cutoff, glynn=True):
    """thewalrus/loop_hafnian_batch.py:260-304."""
    n = A.shape[0]
    assert A.shape[1] == n and D.shape == (n,) and len(fixed_reps) == n - 1
    nz = np.nonzero(list(fixed_reps) + [1])[0]
    Anz, Dnz = A[np.ix_(nz, nz)], D[nz]
    fr = np.asarray(fixed_reps)[nz[:-1]]
    fixed_edges, fixed_m_reps, odd = matched_reps(fr)
    if odd is None:
        edges = _batch_edges_even(fixed_edges)
        Ax = Anz[np.ix_(edges, edges)].astype(np.complex128)
        Dx = Dnz[edges].astype(np.complex128)
        return _batch_core(Ax, Dx, fixed_m_reps, N_cutoff // 2, N_cutoff % 2, glynn, False)
    edges = _batch_edges_odd(fixed_edges, odd)
    Ax = Anz[np.ix_(edges, edges)].astype(np.complex128)
    Dx = Dnz[edges].astype(np.complex128)
    return _batch_core(Ax, Dx, fixed_m_reps, (N_cutoff - 1) // 2, 1 - (N_cutoff % 2), glynn, True)


def loop_hafnian_batch_gamma(A, D, fixed_reps, N_cutoff, glynn=True):
    """thewalrus/loop_hafnian_batch_gamma.py:222-269.  The reference's two sweeps (:52-125, :129-220) differ from
    loop_hafnian_batch only by the loop over the rows of D (XD_S, D_S, oddloop[k], oddloop0[k] per row; AX_S,
    oddVX_S shared), so row k is the batch sweep with loop vector D[k]."""
    n = A.shape[0]
    assert A.shape[1] == n and D.shape[1] == n and len(fixed_reps) == n - 1
    return np.array([loop_hafnian_batch(A, D[k], fixed_reps, N_cutoff, glynn) for k in range(D.shape[0])])


# --------------------------------------------------------------------------------------------------
# permanent (thewalrus/_permanent.py)
# --------------------------------------------------------------------------------------------------


def perm_bbfg(M, k0=0, k1=None, scale=True):
    """thewalrus/_permanent.py:130-168 over Gray-code steps [k0, k1) (step k <-> bin_index k + 1)."""
    M = np.asarray(M)
    n = len(M)
    steps = 2 ** (n - 1)
    k1 = steps if k1 is None else k1
    gray = k0 ^ (k0 >> 1)
    delta = np.array([-1 if (gray >> r) & 1 else 1 for r in range(n)])
    comb = (delta[:, None] * M).sum(axis=0)
    total = 0
    for k in range(k0, k1):
        total = total + (-1 if k & 1 else 1) * np.prod(comb)
        new = (k + 1) ^ ((k + 1) >> 1)
        row = (gray ^ new).bit_length() - 1
        if row < n:
            comb = comb + M[row] * (-2 if new > gray else 2)
        gray = new
    return total / steps if scale else total


def perm_ryser(M, k0=0, k1=None):
    """thewalrus/_permanent.py:86-127 over Gray-code steps [k0, k1)."""
    M = np.asarray(M)
    n = len(M)
    steps = 2**n
    k1 = steps if k1 is None else k1
    gray = k0 ^ (k0 >> 1)
    comb = -sum((M[r] for r in range(n) if (gray >> r) & 1), np.zeros(n, dtype=M.dtype))
    total = 0
    for k in range(k0, k1):
        total = total + (-1 if k & 1 else 1) * np.prod(comb)
        new = (k + 1) ^ ((k + 1) >> 1)
        row = (gray ^ new).bit_length() - 1
        if row < n:
            comb = comb + M[row] * (-1 if new > gray else 1)
        gray = new
    return total


def perm(A, method="bbfg"):
    """thewalrus/_permanent.py:34-83 (closed forms for n <= 3, then Gray-code sums)."""
    n = A.shape[0]
    if n == 0:
        return A.dtype.type(1.0)
    if n == 1:
        return A[0, 0]
    if n == 2:
        return A[0, 0] * A[1, 1] + A[0, 1] * A[1, 0]
    return perm_bbfg(A) if method in ("bbfg", "glynn") else perm_ryser(A)


# --------------------------------------------------------------------------------------------------
# torontonian (thewalrus/_torontonian.py)
# --------------------------------------------------------------------------------------------------


def tor_direct(O, j0=0, j1=None):
    """thewalrus/_torontonian.py:123-154 (numba_tor) over subset indices [j0, j1); bit i of the
    MSB-first index selects mode i, as find_kept_edges(j, ones) does."""
    N = O.shape[0] // 2
    j1 = 2**N if j1 is None else j1
    total = 0.0
    for j in range(j0, j1):
        modes = [i for i in range(N) if (j >> (N - 1 - i)) & 1]
        k = len(modes)
        rows = modes + [i + N for i in modes]
        sub = np.eye(2 * k, dtype=O.dtype) - O[np.ix_(rows, rows)]
        det = np.linalg.det(sub).real if k else 1.0
        total += (-1.0) ** ((N - k) % 2) / np.sqrt(det)
    return total


def tor_recursive(O):
    """thewalrus/_torontonian.py:157-247 — Cholesky-updating recursion over deleted modes."""
    n = O.shape[0] >> 1
    Z = np.empty(2 * n, dtype=int)
    Z[0::2] = np.arange(n)
    Z[1::2] = np.arange(n, 2 * n)
    A = O[np.ix_(Z, Z)]
    L = np.linalg.cholesky(np.eye(2 * n) - A)

    def quad_cholesky(Lp, Zs, idx, mat):
        Ls = Lp[np.ix_(Zs, Zs)].copy()
        for i in range(idx, len(mat)):
            for j in range(idx, i):
                z = Ls[i, :j] @ Ls[j, :j].conj()
                Ls[i, j] = (mat[i, j] - z) / Ls[j, j]
            z = Ls[i, :i] @ Ls[i, :i].conj()
            Ls[i, i] = np.real(np.sqrt(mat[i, i] - z))
        return Ls

    def rec(Lc, modes, Ac):
        tot = 0.0
        start = modes[-1] + 1 if modes else 0
        for i in range(start, n):
            nm = len(Ac) >> 1
            idx = (i - len(modes)) * 2
            Zs = np.concatenate((np.arange(idx), np.arange(idx + 2, 2 * nm)))
            Az = Ac[np.ix_(Zs, Zs)]
            Ls = quad_cholesky(Lc, Zs, idx, np.eye(2 * (nm - 1)) - Az)
            det = np.square(np.prod(np.diag(Ls)))
            tot += (-1) ** (len(modes) + 1) / np.sqrt(det) + rec(Ls, modes + [i], Az)
        return tot

    det = np.square(np.prod(np.diag(L)))
    return np.real(1 / np.sqrt(det) + rec(L, [], A))


def tor(O, recursive=True):
    """thewalrus/_torontonian.py:47-58."""
    if O.shape[0] == 0:
        return 1.0
    return tor_recursive(O) if recursive else np.real(tor_direct(O))


# --------------------------------------------------------------------------------------------------
# SURVEY 8(f) "next" components: Bristolian, loop torontonian, montrealer
# --------------------------------------------------------------------------------------------------


def _subset_rows(j, m):
    """Indices of the set bits of the MSB-first m-bit index j (find_kept_edges(j, ones), _hafnian.py:162-180)."""
    return [i for i in range(m) if (j >> (m - 1 - i)) & 1]


def brs(A, E, j0=0, j1=None):
    """Bristolian, thewalrus/_permanent.py:198-223: sum over row subsets Y of (-1)^(m-|Y|) perm(A_Y^H A_Y + E)
    over subset indices [j0, j1)."""
    A = np.asarray(A, dtype=np.complex128)
    m = A.shape[0]
    j1 = 2**m if j1 is None else j1
    total = 0j
    for j in range(j0, j1):
        rows = _subset_rows(j, m)
        Ay = A[rows, :]
        total += (-1) ** ((m - len(rows)) % 2) * perm_bbfg(Ay.conj().T @ Ay + E)
    return total


def ubrs(A):
    """Unitary Bristolian, thewalrus/_permanent.py:226-249: the same sum without E and without the empty set."""
    A = np.asarray(A, dtype=np.complex128)
    n = A.shape[1]
    return brs(A, np.zeros((n, n)), j0=1)


def fock_threshold_prob(n, d, T):
    """thewalrus/_permanent.py:282-321."""
    n, d = np.array(n), np.array(d)
    fac = float(np.prod([math.factorial(int(x)) for x in n]))
    in_modes = np.array([i for i, c in enumerate(n) for _ in range(int(c))], dtype=int)
    C = np.where(d > 0)[0]
    A = T[np.ix_(C, in_modes)]
    E = np.eye(T.shape[1]) - T.conj().T @ T
    if np.allclose(E, 0):
        return ubrs(A).real / fac
    return brs(A, E[np.ix_(in_modes, in_modes)]).real / fac


def ltor_direct(O, gamma, j0=0, j1=None):
    """Loop torontonian, thewalrus/_torontonian.py:369-412 (numba_ltor) over subset indices [j0, j1):
    sum_S (-1)^(N-|S|) exp(gamma_S (I - O_S)^-1 gamma_S^* / 2) / sqrt(Re det(I - O_S))."""
    O = np.asarray(O, dtype=np.complex128)
    gamma = np.asarray(gamma, dtype=np.complex128)
    N = O.shape[0] // 2
    j1 = 2**N if j1 is None else j1
    total = 0j
    for j in range(j0, j1):
        modes = _subset_rows(j, N)
        k = len(modes)
        rows = modes + [i + N for i in modes]
        sub = np.eye(2 * k) - O[np.ix_(rows, rows)]
        g = gamma[rows]
        top = np.exp(0.5 * g @ np.linalg.solve(sub, g.conj())) if k else 1.0
        det = np.linalg.det(sub).real if k else 1.0
        total += (-1.0) ** ((N - k) % 2) * top / np.sqrt(det)
    return total


def _qmat(cov, hbar=2):
    """thewalrus/quantum/conversions.py:70-96."""
    N = len(cov) // 2
    I = np.identity(N)
    x, xp, p = cov[:N, :N] * 2 / hbar, cov[:N, N:] * 2 / hbar, cov[N:, N:] * 2 / hbar
    aidaj = (x + p + 1j * (xp - xp.T) - 2 * I) / 4
    aiaj = (x - p + 1j * (xp + xp.T)) / 4
    return np.block([[aidaj, aiaj.conj()], [aiaj, aidaj.conj()]]) + np.identity(2 * N)


def threshold_detection_prob(mu, cov, det_pattern, hbar=2, atol=1e-10, rtol=1e-10):
    """thewalrus/_torontonian.py:77-120."""
    mu, cov, det = np.asarray(mu, dtype=float), np.asarray(cov, dtype=float), np.asarray(det_pattern)
    n = cov.shape[0] // 2
    clicked = [i for i in range(n) if det[i] == 1]
    rows = clicked + [i + n for i in clicked]
    if np.allclose(mu, 0, atol=atol, rtol=rtol):
        Q = _qmat(cov, hbar)
        O = np.eye(2 * n) - np.linalg.inv(Q)
        Os = O[np.ix_(rows, rows)]
        t = tor_direct(Os) if len(clicked) else 1.0
        return (t / np.sqrt(np.linalg.det(Q))).real
    alpha = np.concatenate((mu[:n] + 1j * mu[n:], mu[:n] - 1j * mu[n:])) / np.sqrt(2 * hbar)
    sigma = _qmat(cov, hbar).conj()
    inv_sigma = np.linalg.inv(sigma)
    O = np.eye(2 * n) - inv_sigma
    gamma = (inv_sigma @ alpha).conj()
    vac = (np.exp(-0.5 * alpha.conj() @ inv_sigma @ alpha).real / np.sqrt(np.linalg.det(sigma))).real
    lt = ltor_direct(O[np.ix_(rows, rows)], gamma[rows]) if len(clicked) else 1.0
    return float(vac * np.real(lt))


def montrealer(Sigma, zeta=None, j0=1, j1=None):
    """thewalrus/_montrealer.py:37-57 (and :78-102 with zeta): over subset labels p in [j0, j1) of 2^n,
    (-1)^(n+1) [ sum_p (-1)^(|p|+1) tr(Sigma_p^n) / (2n) + sum_p (-1)^(|p|+1) conj(zeta_p) Sigma_p^(n-1) zeta_p / 2 ].
    dec2bin (:17-34) selects mode i when bit i of the MSB-first n-bit label is set."""
    Sigma = np.asarray(Sigma, dtype=np.complex128)
    n = len(Sigma) // 2
    j1 = 2**n if j1 is None else j1
    val, val_loops = 0j, 0j
    for p in range(max(j0, 1), j1):
        modes = _subset_rows(p, n)
        pos = modes + [i + n for i in modes]
        sub = Sigma[np.ix_(pos, pos)]
        sign = (-1) ** (len(modes) + 1)
        val += sign * np.trace(np.linalg.matrix_power(sub, n))
        if zeta is not None:
            z = np.asarray(zeta, dtype=np.complex128)[pos]
            val_loops += sign * (z.conj() @ np.linalg.matrix_power(sub, n - 1) @ z)
    return (-1) ** (n + 1) * (val / (2 * n) + val_loops / 2)


def mtl(A):
    """thewalrus/_montrealer.py:121-135: montrealer of Xmat(n) @ A."""
    A = np.asarray(A)
    n = len(A) // 2
    X = np.block([[np.zeros((n, n)), np.eye(n)], [np.eye(n), np.zeros((n, n))]])
    return montrealer(X @ A)


def lmtl(A, zeta):
    """thewalrus/_montrealer.py:105-118."""
    A = np.asarray(A)
    n = len(A) // 2
    X = np.block([[np.zeros((n, n)), np.eye(n)], [np.eye(n), np.zeros((n, n))]])
    return montrealer(X @ A, zeta)


# --------------------------------------------------------------------------------------------------
# pure definitions used as exact cross-checks (thewalrus/reference.py:195-284)
# --------------------------------------------------------------------------------------------------


def hafnian_by_matchings(A, loop=False):
    """Sum over perfect matchings (loop: single-pair matchings) — exact on Python ints."""
    n = len(A)

    def rec(rem):
        if not rem:
            return 1
        i = rem[0]
        tot = 0
        if loop:
            tot += A[i][i] * rec(rem[1:])
        for idx in range(1, len(rem)):
            j = rem[idx]
            tot += A[i][j] * rec(rem[1:idx] + rem[idx + 1:])
        return tot

    if n % 2 and not loop:
        return 0
    return rec(tuple(range(n)))
