"""Full-size goldens for the BASELINE configs (VERDICT r01 "next round" item 1).

Run once in the authoring container (`python tests/golden/make_golden_fullsize.py [stage ...]`): it imports the
REFERENCE (/root/reference, `dask` stubbed) and the long-double C oracle, and writes
tests/golden/reference_fullsize.json.  The GPU box has neither, so the JSON is committed.  Stages are
independent and the file is updated after each one, so slow stages can run in the background.

Every entry stores a fingerprint of its input (sum and sum of squares of the matrix) so that the GPU-side test can
prove it regenerated the same input from the same seed (inputs come from bench.make_input / the recipes below).

Stages
  structured  full-size inputs whose value factorises into pieces the LONG-DOUBLE oracle finishes in seconds:
              haf([[0,B],[B^T,0]]) = perm(B) at n = 50 / 56, haf(P (A1 (+) A2) P^T) = haf(A1) haf(A2) at n = 50 / 56 / 64,
              perm(P (B1 (+) B2) Q) = perm(B1) perm(B2) at n = 32 / 40 — the kernels do the full 2^24 .. 2^39 step
              sweep whatever the entries are, so a grid-stride hole, a mis-sharded tail or a 2^31 index overflow shows.
  exact       integer perfect-matching counts from thewalrus.reference.hafnian (reference.py:229-284) and int64
              recursive_hafnian (_hafnian.py:978-1045).
  c1          hafnian24 (BASELINE config 1): reference numba + long-double oracle.
  perm32      BASELINE config 2 on the bench input: long-double oracle over all 2^31 steps (+ reference `perm`).
  tor48       BASELINE config 4 on the bench input: reference rec_torontonian + C port (double and long double).
  gbs16       BASELINE config 3: all 10^5 bench patterns through the C port (double and long double; sum and every
              probability), 200 sampled patterns through the reference's density_matrix_element.
  hafnian50   the metric's input: full 2^24-subset sum through the double C oracle (Kahan), the reference numba
              `hafnian` itself, and 64 long-double windows as a yardstick for the double results.
"""
import json
import os
import sys
import time
import types

import numpy as np

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")     # numba prange x threaded OpenBLAS oversubscribes at n >= 50
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, "reference_fullsize.json")

from oracle import c_oracle as co  # noqa: E402
import bench  # noqa: E402


def ref():
    _d = types.ModuleType("dask")
    _d.delayed = lambda f, *a, **k: f
    _d.compute = lambda *a, **k: a
    sys.modules.setdefault("dask", _d)
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    import thewalrus

    return thewalrus


def enc(z):
    z = complex(z)
    return {"re": z.real, "im": z.imag}


def fp(*arrays):
    """Input fingerprint: [sum, sum of squares] over all arrays (complex -> re, im pairs)."""
    s = sum(complex(np.sum(np.asarray(a))) for a in arrays)
    q = sum(complex(np.sum(np.asarray(a).astype(np.complex128) ** 2)) for a in arrays)
    return [s.real, s.imag, q.real, q.imag]


def load():
    if os.path.exists(OUT):
        with open(OUT) as fh:
            return json.load(fh)
    return {}


def save(d):
    tmp = OUT + ".tmp"
    with open(tmp, "w") as fh:
        json.dump(d, fh, indent=1)
    os.replace(tmp, OUT)


# ---- input recipes shared with tests/test_gpu_fullsize.py (imported from here by the test) -----------------------------------
def bipartite_input(n, seed):
    """A = [[0, B], [B^T, 0]] with B an (n/2 x n/2) complex Gaussian matrix scaled by 1/sqrt(n/2), vertices shuffled."""
    h = n // 2
    rng = np.random.default_rng(seed)
    B = (rng.standard_normal((h, h)) + 1j * rng.standard_normal((h, h))) / np.sqrt(h)
    A = np.zeros((n, n), dtype=np.complex128)
    A[:h, h:] = B
    A[h:, :h] = B.T
    p = rng.permutation(n)
    return B, np.ascontiguousarray(A[np.ix_(p, p)])


def direct_sum_input(n1, n2, seed):
    rng = np.random.default_rng(seed)
    parts = []
    for k in (n1, n2):
        G = rng.standard_normal((k, k)) + 1j * rng.standard_normal((k, k))
        parts.append((G + G.T) / np.sqrt(k))
    n = n1 + n2
    A = np.zeros((n, n), dtype=np.complex128)
    A[:n1, :n1] = parts[0]
    A[n1:, n1:] = parts[1]
    p = rng.permutation(n)
    return parts, np.ascontiguousarray(A[np.ix_(p, p)])


def block_perm_input(n1, n2, seed):
    """P (B1 (+) B2) Q with B_i blocks of Haar unitaries: perm = perm(B1) perm(B2)."""
    rng = np.random.default_rng(seed)
    parts = []
    for k in (n1, n2):
        Z = (rng.standard_normal((2 * k, 2 * k)) + 1j * rng.standard_normal((2 * k, 2 * k))) / np.sqrt(2)
        Q, R = np.linalg.qr(Z)
        parts.append(np.ascontiguousarray((Q * (np.diag(R) / np.abs(np.diag(R))))[:k, :k]) * 2.0)
    n = n1 + n2
    M = np.zeros((n, n), dtype=np.complex128)
    M[:n1, :n1] = parts[0]
    M[n1:, n1:] = parts[1]
    return parts, np.ascontiguousarray(M[np.ix_(rng.permutation(n), rng.permutation(n))])


def er_graph(n, p, seed):
    rng = np.random.default_rng(seed)
    G = (rng.random((n, n)) < p).astype(np.int64)
    A = np.triu(G, 1)
    return A + A.T


# ---- stages -------------------------------------------------------------------------------------------------------------------
def stage_structured(d):
    out = []
    for n, seed in ((50, 915050), (56, 915056)):
        B, A = bipartite_input(n, seed)
        out.append({"kind": "haf_bipartite", "n": n, "seed": seed, "fp": fp(A), "value": enc(co.perm(B, "bbfg", 0, True)),
                    "how": f"perm of the {n // 2}x{n // 2} block, long-double Glynn oracle"})
        print(out[-1]["kind"], n, out[-1]["value"], flush=True)
    for n1, n2, seed in ((24, 26, 925050), (28, 28, 925056), (30, 34, 925064)):
        parts, A = direct_sum_input(n1, n2, seed)
        v = co.hafnian(parts[0], False, 0, True) * co.hafnian(parts[1], False, 0, True)
        lv = co.hafnian(parts[0], True, 0, True) * co.hafnian(parts[1], True, 0, True)
        out.append({"kind": "haf_direct_sum", "n": n1 + n2, "n1": n1, "n2": n2, "seed": seed, "fp": fp(A), "value": enc(v),
                    "loop_value": enc(lv), "how": "product of the two blocks' (loop) hafnians, long-double oracle"})
        print(out[-1]["kind"], n1 + n2, out[-1]["value"], flush=True)
    for n1, n2, seed in ((16, 16, 935032), (20, 20, 935040), (18, 22, 935140)):
        parts, M = block_perm_input(n1, n2, seed)
        v = co.perm(parts[0], "bbfg", 0, True) * co.perm(parts[1], "bbfg", 0, True)
        out.append({"kind": "perm_blocks", "n": n1 + n2, "n1": n1, "n2": n2, "seed": seed, "fp": fp(M), "value": enc(v),
                    "how": "product of the two blocks' permanents, long-double Glynn oracle"})
        print(out[-1]["kind"], n1 + n2, out[-1]["value"], flush=True)
    d["structured"] = out


def stage_exact(d):
    tw = ref()
    from thewalrus import reference as twref
    from thewalrus._hafnian import recursive_hafnian

    out = []
    for n, p in ((8, 0.7), (10, 0.6), (12, 0.6), (14, 0.5)):
        A = er_graph(n, p, 940000 + n)
        L = A + np.diag((np.arange(n) % 2).astype(np.int64))
        out.append({"n": n, "p": p, "seed": 940000 + n, "fp": fp(A), "hafnian": int(twref.hafnian(A)),
                    "loop_hafnian": int(twref.hafnian(L, loop=True)), "source": "thewalrus.reference.hafnian (reference.py:229-284)"})
        assert out[-1]["hafnian"] == int(recursive_hafnian(A))
        print("exact", n, out[-1]["hafnian"], out[-1]["loop_hafnian"], flush=True)
    for n, p in ((20, 0.5), (24, 0.45), (28, 0.4), (32, 0.3), (36, 0.25), (40, 0.2)):
        A = er_graph(n, p, 940000 + n)
        v = int(recursive_hafnian(A))
        assert 0 <= v < 2 ** 52, (n, v)
        out.append({"n": n, "p": p, "seed": 940000 + n, "fp": fp(A), "hafnian": v,
                    "source": "int64 recursive_hafnian (_hafnian.py:978-1045)"})
        print("exact", n, v, flush=True)
    d["exact"] = out
    del tw


def stage_c1(d):
    tw = ref()
    _, n, A = bench.make_input("hafnian24")
    d["hafnian24"] = {"fp": fp(A), "reference": enc(tw.hafnian(A)), "reference_loop": enc(tw.hafnian(A, loop=True)),
                      "oracle_ld": enc(co.hafnian(A, False, 0, True)), "oracle_ld_loop": enc(co.hafnian(A, True, 0, True))}
    print("hafnian24", d["hafnian24"], flush=True)


def stage_perm32(d):
    _, n, U = bench.make_input("perm32")
    t0 = time.time()
    v = co.perm(U, "bbfg", 0, True)
    e = {"fp": fp(U), "oracle_ld": enc(v), "oracle_ld_seconds": time.time() - t0,
         "oracle_double": enc(co.perm(U, "bbfg", 0, False))}
    print("perm32 ld", e, flush=True)
    d["perm32"] = e
    save(d)
    tw = ref()
    t0 = time.time()
    e["reference_bbfg"] = enc(tw.perm(U, method="bbfg"))
    e["reference_seconds"] = time.time() - t0
    print("perm32 ref", e["reference_bbfg"], e["reference_seconds"], flush=True)


def stage_tor48(d):
    tw = ref()
    from thewalrus._torontonian import rec_torontonian

    _, n, O = bench.make_input("tor48")
    e = {"fp": fp(O)}
    t0 = time.time()
    e["oracle_ld"] = co.tor_recursive(O, True)
    e["oracle_double"] = co.tor_recursive(O, False)
    print("tor48 oracle", e, time.time() - t0, flush=True)
    t0 = time.time()
    e["reference_rec"] = enc(rec_torontonian(O))
    e["reference_seconds"] = time.time() - t0
    _, _, (Ol, gam) = bench.make_input("ltor48")
    e["ltor_fp"] = fp(Ol, gam)
    e["ltor_oracle_ld"] = enc(co.ltor_direct(Ol, gam, 0, None, 0, True))
    e["ltor_oracle_double"] = enc(co.ltor_direct(Ol, gam, 0, None, 0, False))
    d["tor48"] = e
    print("tor48", e, flush=True)
    del tw


def stage_gbs16(d):
    tw = ref()
    from thewalrus.quantum import density_matrix_element

    M, mu, cov, pats, A, gamma, rpt = bench.gbs_inputs("gbs16", 100000)
    import thewalrus_b200.quantum as wq

    pref = wq._prefactor(mu, cov)
    scale = np.exp(-wq._log_factorial_sums(pats))
    e = {"fp": fp(mu, cov, pats), "B": int(len(pats))}
    for name, ld in (("double", False), ("ld", True)):
        t0 = time.time()
        lh = co.lhaf_patterns(A, gamma, rpt, True, 0, ld)
        p = (lh * pref).real * scale
        e["oracle_%s_sum" % name] = float(np.sum(np.sort(p)))
        e["oracle_%s_seconds" % name] = time.time() - t0
        if ld:
            np.save(os.path.join(HERE, "gbs16_probabilities_ld.npy"), p.astype(np.float64))
        print("gbs16", name, e["oracle_%s_sum" % name], time.time() - t0, flush=True)
        d["gbs16"] = e
        save(d)
    order = np.argsort(pats.sum(axis=1), kind="stable")
    pick = sorted(set(int(order[int(i)]) for i in np.linspace(0, len(pats) - 1, 200)))
    t0 = time.time()
    e["sample_index"] = pick
    e["sample_reference"] = [float(density_matrix_element(mu, cov, list(pats[i]), list(pats[i])).real) for i in pick]
    e["sample_reference_seconds"] = time.time() - t0
    print("gbs16 reference sample", time.time() - t0, flush=True)
    del tw


def stage_hafnian50(d):
    _, n, A = bench.make_input("hafnian50")
    x = co.matched_order(A)
    Ax = np.ascontiguousarray(A[np.ix_(x, x)])
    e = d.get("hafnian50", {"fp": fp(A)})
    total = 1 << (n // 2 - 1)
    if "windows_ld" not in e:
        rng = np.random.default_rng(50)
        starts = sorted(int(s) for s in rng.integers(0, total - 64, 64))
        e["windows_ld"] = [{"j0": s, "j1": s + 32, "ld": enc(co.hafnian_range(Ax, s, s + 32, None, 0, True)),
                            "double": enc(co.hafnian_range(Ax, s, s + 32, None, 0, False))} for s in starts]
        d["hafnian50"] = e
        save(d)
    if "oracle_double" not in e:
        t0 = time.time()
        # 64 chunks, each a Kahan sum inside the oracle; chunk partials kept so a sharding bug can be localised
        chunks = [co.hafnian_range(Ax, c * (total // 64), (c + 1) * (total // 64), None, 0, False) for c in range(64)]
        acc = sum(chunks[1:], chunks[0]) * 0.5 ** (n // 2 - 1)
        e["oracle_double"] = enc(acc)
        e["oracle_double_chunks"] = [enc(c) for c in chunks]
        e["oracle_double_seconds"] = time.time() - t0
        print("hafnian50 oracle double", e["oracle_double"], time.time() - t0, flush=True)
        d["hafnian50"] = e
        save(d)
    if "reference" not in e:
        tw = ref()
        G = np.random.default_rng(1).standard_normal((12, 12))
        tw.hafnian(G + G.T)     # JIT warm-up
        t0 = time.time()
        e["reference"] = enc(tw.hafnian(A))
        e["reference_seconds"] = time.time() - t0
        print("hafnian50 reference", e["reference"], time.time() - t0, flush=True)
        d["hafnian50"] = e
        save(d)


STAGES = {"structured": stage_structured, "exact": stage_exact, "c1": stage_c1, "perm32": stage_perm32,
          "tor48": stage_tor48, "gbs16": stage_gbs16, "hafnian50": stage_hafnian50}

if __name__ == "__main__":
    todo = sys.argv[1:] or list(STAGES)
    for name in todo:
        d = load()
        t0 = time.time()
        STAGES[name](d)
        save(d)
        print(f"stage {name} done in {time.time() - t0:.1f} s", flush=True)
