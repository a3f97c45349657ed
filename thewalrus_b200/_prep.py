"""Host-side preparation shared by the hafnian entry points (integer bookkeeping only).

``matched_reps`` mirrors thewalrus/_hafnian.py:80-159: greedily pair vertices into "edges" with
repetition counts so that the subset sum runs over prod(edge_reps + 1) terms.
"""
import numpy as np


def matched_reps(reps):
    """Pair up repeated vertices.

    Returns ``(x, edge_reps, oddmode)``: ``x`` has length ``2 * n_edges`` and vertex ``x[i]`` is paired
    with ``x[i + n_edges]``; ``edge_reps[i]`` is how often that pair is repeated; ``oddmode`` is the one
    left-over vertex or ``None``.  Same greedy rule and tie-breaking as the reference
    (thewalrus/_hafnian.py:97-159): repeatedly sort by (reps, index) descending; if the largest count
    exceeds twice the runner-up, pair the vertex with itself, else pair the two largest.
    """
    n = len(reps)
    if sum(reps) == 0:
        return np.array([], dtype=np.int64), np.array([], dtype=np.int64), None
    if n > 1 and all(r == 1 for r in reps):
        # closed form of the greedy rule for all-ones repetitions (the plain hafnian / loop hafnian call):
        # vertices n-1, n-3, ... are paired with n-2, n-4, ...; an odd n leaves vertex 0 over
        first = np.arange(n - 1, 0, -2, dtype=np.int64)
        return np.concatenate([first, first - 1]), np.ones(len(first), dtype=np.int64), (0 if n % 2 else None)

    pool = [(int(r), i) for i, r in zip(range(n), reps) if r > 0]
    first, second, counts = [], [], []
    while len(pool) > 1 or (len(pool) == 1 and pool[0][0] > 1):
        pool.sort(reverse=True)
        (r0, v0) = pool[0]
        if len(pool) == 1 or r0 > 2 * pool[1][0]:
            first.append(v0)
            second.append(v0)
            counts.append(r0 // 2)
            if r0 % 2 == 0:
                pool = pool[1:]
            else:
                pool[0] = (1, v0)
        else:
            (r1, v1) = pool[1]
            first.append(v0)
            second.append(v1)
            counts.append(r1)
            if r0 > r1:
                pool = [(r0 - r1, v0)] + pool[2:]
            else:
                pool = pool[2:]
    oddmode = pool[0][1] if len(pool) == 1 else None
    x = np.asarray(first + second, dtype=np.int64)
    return x, np.asarray(counts, dtype=np.int64), oddmode


def glynn_steps(edge_reps, glynn=True, has_odd=False):
    """Number of subset indices (reference ``steps``; thewalrus/_hafnian.py:432-435, 535-538)."""
    edge_reps = [int(e) for e in edge_reps]
    if len(edge_reps) == 0:
        return 1
    if glynn and not has_odd:
        s = (edge_reps[0] + 2) // 2
        rest = edge_reps[1:]
    else:
        s = 1
        rest = edge_reps
    for e in rest:
        s *= e + 1
    return s


def shard_range(total, rank, world):
    """Contiguous shard [lo, hi) of ``range(total)`` owned by ``rank`` of ``world``."""
    return (rank * total) // world, ((rank + 1) * total) // world


def dd_sum(pairs):
    """Sum (hi, lo) pairs in the given (fixed) order with an error-free two-sum; returns float."""
    hi, lo = 0.0, 0.0
    for (h, l) in pairs:
        s = hi + h
        bb = s - hi
        e = (hi - (s - bb)) + (h - bb)
        e += lo + l
        hi2 = s + e
        lo = e - (hi2 - s)
        hi = hi2
    return hi, lo
