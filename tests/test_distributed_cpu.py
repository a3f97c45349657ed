"""N>1 path on CPU: two gloo ranks shard the subset index, each evaluates its contiguous range (with the
oracle standing in for the kernel), and one all-reduce + fixed-order compensated sum reproduces the
single-process result bit-for-bit on every rank."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from oracle import c_oracle as co
    from thewalrus_b200 import _engine

    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(4)
    n = 14
    G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    A = G + G.T
    x = co.matched_order(A)
    Ax = A[np.ix_(x, x)]
    steps = 1 << (n // 2 - 1)

    def runner(lo, hi):
        v = co.hafnian_range(Ax, lo, hi)
        return np.array([v.real, 0.0, v.imag, 0.0])

    table = _engine.run_sharded(steps, runner, group=True)
    total = _engine.combine4(table) * 0.5 ** (n // 2 - 1)
    # integer (exact) path used by perm on int64 input
    part = np.array([float(rank + 1), float(7 * rank)])
    t2 = _engine.allreduce_partials(part, None)
    # batched front end: PATTERNS are sharded (ragged: 7 patterns over 2 ranks) and all-gathered
    from oracle import walrus_oracle as wo

    A5 = A[:5, :5]
    g5 = np.diag(A)[:5].copy()
    rpt = np.random.default_rng(5).integers(0, 3, (7, 5)).astype(np.int32)

    def local(rows):
        return np.array([wo.loop_hafnian(A5, g5, [int(x) for x in r]) for r in rows], dtype=np.complex128)

    pats = _engine.run_sharded_patterns(A5, g5, rpt, True, True, None, local=local)

    # batched-matrix / multi-gamma form: the per-problem index arrays must be sliced with the shard.  The GPU
    # evaluator of a rank is replaced by the oracle (this box has no GPU); the sharding code is what runs here.
    def oracle_local(Ast, gam, r, glynn=True, device=None, want_ms=False, gamma_index=None, A_index=None):
        return np.array([wo.loop_hafnian(Ast[a], gam[g], [int(x) for x in row])
                         for row, g, a in zip(r, gamma_index, A_index)], dtype=np.complex128)

    _engine.lhaf_patterns_local = oracle_local
    Ast = np.stack([A5, 0.5 * A5, A5 * (1 + 0.1j)])
    gam = np.stack([g5, 2 * g5])
    gi = np.array([0, 1, 1, 0, 1, 0, 0], dtype=np.int32)
    ai = np.array([2, 0, 1, 1, 0, 2, 1], dtype=np.int32)
    pats2 = _engine.run_sharded_patterns(Ast, gam, rpt, True, True, None, gamma_index=gi, A_index=ai)
    q.put((rank, total, table.tolist(), t2.tolist(), pats.tolist(), pats2.tolist()))
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_sum():
    sys.path.insert(0, ROOT)
    from oracle import c_oracle as co

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(4)
    n = 14
    G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    A = G + G.T
    want = co.hafnian(A)
    (r0, v0, tab0, i0, p0, m0), (r1, v1, tab1, i1, p1, m1) = res
    from oracle import walrus_oracle as wo

    rpt = np.random.default_rng(5).integers(0, 3, (7, 5)).astype(np.int32)
    wantp = [wo.loop_hafnian(A[:5, :5], np.diag(A)[:5].copy(), [int(x) for x in r]) for r in rpt]
    assert p0 == p1 and np.allclose(np.array(p0), np.array(wantp), rtol=1e-13, atol=0)
    A5, g5 = A[:5, :5], np.diag(A)[:5].copy()
    Ast, gam = [A5, 0.5 * A5, A5 * (1 + 0.1j)], [g5, 2 * g5]
    gi, ai = [0, 1, 1, 0, 1, 0, 0], [2, 0, 1, 1, 0, 2, 1]
    wantm = [wo.loop_hafnian(Ast[a], gam[g], [int(x) for x in r]) for r, g, a in zip(rpt, gi, ai)]
    assert m0 == m1 and np.allclose(np.array(m0), np.array(wantm), rtol=1e-13, atol=0)
    assert v0 == v1, "ranks must agree bit-for-bit"
    assert tab0 == tab1 and len(tab0) == world
    assert abs(v0 - want) / abs(want) < 1e-12
    assert i0 == i1 == [[1.0, 0.0], [2.0, 7.0]]
