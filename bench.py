#!/usr/bin/env python
"""Headline benchmark: BASELINE.json metric "hafnian n=50 complex128 subsets/s & wall-time at 1/2/4/8 B200;
% FP64 peak".

One "step" = one complete hafnian of a random 50x50 complex128 symmetric matrix (2^24 Glynn subsets,
seed 1000*1+50 as in SURVEY.md 8d).  With N GPUs the subset index is sharded in contiguous ranges (strong
scaling: total work fixed) and the partial sums are combined by one all-reduce inside the timed region.

  python bench.py --gpus N --steps K --warmup W                   # GPU arm (torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K --warmup W  # CPU arm: the reference's numba path from
                                                                  # baseline/_ref on the host cores (C port if absent)

The default N = 1 run also measures the other BASELINE configs (hafnian24, perm32, tor48, gbs16) and prints them in a
"secondary" object of the same JSON line, each with value / e2e / roofline / clocks / cpu_baseline and the error of
the complete result against the offline golden (tests/golden/reference_fullsize.json).

Other workloads: --workload hafnian24|hafnian56|lhaf50|perm32|perm40|tor48|gbs16 and the SURVEY 8(f) components
ltor48, mtl14, brs12, hsample8 (batched chain-rule photon-number sampler; unit samples/s).
"""
import os
import sys

if "reference" in sys.argv:   # the numba reference: one BLAS thread per prange worker (6x cliff otherwise);
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")   # must be set before numpy loads OpenBLAS
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_wb200")

import argparse  # noqa: E402
import ctypes  # noqa: E402
import json  # noqa: E402
import math  # noqa: E402
import statistics  # noqa: E402
import subprocess  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hafnian n=50 complex128 subsets/s"
SECONDARY = ["hafnian24", "perm32", "tor48", "gbs16"]
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the `ncu --set full` captures
# under profiles/ (file named next to each number)
MEASURED_TRAFFIC = {
    # haf_sym_kernel, complete 2^24-subset launch inside bench.py: 19.1 MB read + 94.2 MB written in 4.3 s.  The writes are the
    # dirty lines of the 256 MiB L2-flush fill that precedes every timed launch, evicted while the kernel runs; the same kernel
    # on 2^18 subsets without a flush before it moves 0.5 MB (profiles/r02_ncu_haf50_sym.txt)
    "hafnian50": (19121920 + 94205184, "profiles/r02_launches_hafnian50_final3.csv"),
    "perm32": (79104, "profiles/r01_ncu_perm32_v3.txt"),
    "tor48": (73216, "profiles/r01_ncu_tor48_v3b.txt (round-1 kernel, square storage; the packed kernel of round 2 was not re-captured)"),
}
GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_fullsize.json")


def make_input(workload):
    """Synthetic inputs exactly as SURVEY.md 8(d) defines them (seed = 1000*config + size)."""
    kind = workload.rstrip("0123456789")
    n = int(workload[len(kind):])
    if kind in ("hafnian", "lhaf"):
        rng = np.random.default_rng(1000 * (1 if n <= 50 else 5) + n)
        G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        A = G + G.T
        return kind, n, A
    if kind == "perm":
        rng = np.random.default_rng(1000 * 2 + n)
        Z = (rng.standard_normal((2 * n, 2 * n)) + 1j * rng.standard_normal((2 * n, 2 * n))) / np.sqrt(2)
        Q, R = np.linalg.qr(Z)
        U = Q * (np.diag(R) / np.abs(np.diag(R)))  # Haar phase fix (thewalrus/random.py:115-134)
        return kind, n, np.ascontiguousarray(U[:n, :n])
    if kind in ("tor", "ltor"):  # n = 2N: config 4 of SURVEY 8(d), all N detectors click
        from thewalrus_b200.quantum import Qmat

        N = n // 2
        # squeezing r = 1.5 instead of the 0.5 of config 3: with 0.2 photons per mode the all-click probability of 24
        # modes is ~1e-18 and the alternating sum is pure rounding noise (in the reference too)
        mu, cov, _ = make_gbs_state(N, 1, seed=1000 * 4 + n, r=1.5)
        if kind == "tor":   # zero-mean state: threshold_detection_prob = tor(I - Q^-1) / sqrt(det Q)  (_torontonian.py:98-104)
            return kind, n, np.ascontiguousarray(np.identity(n) - np.linalg.inv(Qmat(cov)))
        sigma_inv = np.linalg.inv(Qmat(cov).conj())   # displaced state: the ltor arguments of _torontonian.py:106-120
        alpha = np.concatenate([mu[:N] + 1j * mu[N:], mu[:N] - 1j * mu[N:]]) / 2.0
        return kind, n, (np.ascontiguousarray(np.identity(n) - sigma_inv), (sigma_inv @ alpha).conj())
    if kind == "mtl":  # n = modes; 2n x 2n complex symmetric adjacency-like matrix
        rng = np.random.default_rng(1000 * 6 + n)
        G = rng.standard_normal((2 * n, 2 * n)) + 1j * rng.standard_normal((2 * n, 2 * n))
        return kind, n, (G + G.T) / np.sqrt(8.0 * n)
    if kind == "brs":  # n x n block of a 2n-mode Haar unitary (lossy interferometer) and E = I - A^H A
        rng = np.random.default_rng(1000 * 7 + n)
        Z = (rng.standard_normal((2 * n, 2 * n)) + 1j * rng.standard_normal((2 * n, 2 * n))) / np.sqrt(2)
        Q, R = np.linalg.qr(Z)
        A = np.ascontiguousarray((Q * (np.diag(R) / np.abs(np.diag(R))))[:n, :n])
        return kind, n, (A, np.identity(n) - A.conj().T @ A)
    if kind == "hsample":  # n = modes: GBS state as in config 3, smaller squeezing so chains stay below the cutoff
        mu, cov, _ = make_gbs_state(n, 1, seed=1000 * 8 + n, r=0.4)
        return kind, n, (mu, cov)
    raise SystemExit(f"unknown workload {workload}")


def make_gbs_state(M, B, seed, r=0.5, eta=0.8, hbar=2.0, mean_photons=0.45, max_total=10):
    """BASELINE config 3: M-mode state  cov = eta (hbar/2) S S^T + (1 - eta)(hbar/2) I  with
    S = interferometer(U_Haar) . squeezing(r), means 0.3 N(0,1), and B patterns drawn i.i.d. Poisson(0.45) per
    mode, rejected if the total exceeds ``max_total``.  Returns (mu, cov, patterns[B, M])."""
    rng = np.random.default_rng(seed)
    Z = (rng.standard_normal((M, M)) + 1j * rng.standard_normal((M, M))) / np.sqrt(2)
    Q, R = np.linalg.qr(Z)
    U = Q * (np.diag(R) / np.abs(np.diag(R)))
    X, Y = U.real, U.imag
    Sint = np.block([[X, -Y], [Y, X]])                     # symplectic of a passive interferometer
    Ssq = np.diag(np.concatenate([np.exp(-r) * np.ones(M), np.exp(r) * np.ones(M)]))
    S = Sint @ Ssq
    cov = eta * (hbar / 2) * S @ S.T + (1 - eta) * (hbar / 2) * np.identity(2 * M)
    mu = 0.3 * rng.standard_normal(2 * M)
    pats = np.zeros((0, M), dtype=np.int32)
    while len(pats) < B:
        cand = rng.poisson(mean_photons, size=(2 * (B - len(pats)) + 16, M)).astype(np.int32)
        pats = np.concatenate([pats, cand[cand.sum(axis=1) <= max_total]])
    return mu, cov, np.ascontiguousarray(pats[:B])


def tor_tree_flops(N, DC=9, aug=0):
    """Executed flops of the Schur-complement tree kernel (torontonian.cu) per torontonian, counted from its
    loops: an included child of a breadth-first node of dimension d updates (d-2)^2 complex entries with four
    a*conj(b) products (6 flops) and four real-scaled subtractions (4 flops) = 40 flops; a 2-mode leaf node
    costs ~70 flops for its 4 subsets; the shared leading-mode eliminations are lower order and included."""
    DC = min(DC, N)
    per_prefix = 0.0
    for lvl in range(max(0, DC - 2)):
        d = 2 * (DC - lvl) + aug
        per_prefix += (1 << lvl) * (d - 2) ** 2 * 40.0
    per_prefix += (1 << max(0, DC - 2)) * (150.0 if aug else 70.0)
    lead = 0.0
    for i in range(N - DC):  # on average half of the leading modes are eliminated, on the full trailing block
        d = 2 * (N - i) + aug
        lead += 0.5 * ((d - 1) ** 2 + (d - 2) ** 2) * 10.0
    prefixes = 1 << (N - DC)
    return prefixes * per_prefix + (prefixes / 32.0) * lead


def haf_sym_entries(n):
    """Useful entries of one n x n product the symmetric-half kernel computes (hafnian_sym.cu; 0 = size not covered): the
    row panel of a vertex pair in tile T (tiles = 4 vertex pairs = 8 rows/columns) computes the columns of tiles >= T
    and of the tail pair; the tail pair's panel only its own 2 x 2 block.  Sizes with n/2 = 2, 3 (mod 4) run zero-padded in
    the next whole-tile shape: only the entries of real vertices are counted."""
    m = n // 2
    if n % 2 or not 36 <= n <= 64:
        return 0
    if m % 4 == 1:                                   # whole tiles + one tail pair
        TF = m // 4
        return sum(8 * (8 * (TF - T) + 2) for T in range(TF)) + 4
    return 4 * sum(1 for i in range(m) for p in range(m) if p // 4 >= i // 4)


def units_and_flops(kind, n):
    """(units per step, executed-algorithm flops per unit of THIS implementation, reference-algorithm flops per
    unit, text of the model, FP64 pipe the kernel issues to)."""
    if kind in ("hafnian", "lhaf"):
        m = n // 2
        nprod = (m - 1) // 2
        if kind == "hafnian" and haf_sym_entries(n) and os.environ.get("WB200_HAF_SYM", "1") != "0":
            # symmetric-half kernel (hafnian_sym.cu): of every product only the tiles on and above the diagonal
            return (1 << (m - 1), 8.0 * n * haf_sym_entries(n) * nprod, 8.0 * n**3 * (m - 1),
                    "EXECUTED useful flops of haf_sym_kernel: 8 n E floor((n/2-1)/2) per subset, E = %d of the n^2 = %d entries of "
                    "a product (B_k = M^(k-1) A' is symmetric: only the 8 x 8 tiles on and above the diagonal are computed); "
                    "trace pairing halves the reference's n/2-1 products.  The row-panel kernel of round 1 executes 8 n^3 "
                    "floor((n/2-1)/2) = %.3g per subset, the reference's algorithm 8 n^3 (n/2-1) = %.3g"
                    % (haf_sym_entries(n), n * n, 8.0 * n**3 * nprod, 8.0 * n**3 * (m - 1)),
                    "FP64 DMMA.8x8x4 (tensor pipe; same flop rate as the FP64 FMA pipe)")
        return (1 << (m - 1), 8.0 * n**3 * nprod, 8.0 * n**3 * (m - 1),
                "8 n^3 floor((n/2-1)/2) per subset: trace pairing halves the reference's 8 n^3 (n/2-1) product chain",
                "FP64 DMMA.8x8x4 (tensor pipe; same flop rate as the FP64 FMA pipe)")
    if kind == "perm":
        return (1 << (n - 1), 8.0 * n - 4, 8.0 * n - 4,
                "8n-4 per Gray-code step (n complex adds + n-1 complex multiplies + accumulate); these occupy 6n-2 "
                "FP64 issue slots, so the flop fraction cannot exceed (8n-4)/(2(6n-2)) ~ 67 % of the DFMA peak",
                "FP64 DFMA/DMUL/DADD (vector pipe)")
    if kind == "tor":
        N = n // 2
        f = tor_tree_flops(N) / float(1 << N)
        # direct method of the reference: complex Cholesky (4/3)(2k)^3 averaged over subsets, E[k^3] over Binomial(N, 1/2)
        ek3 = N * (N - 1) * (N - 2) / 8.0 + 3 * N * (N - 1) / 4.0 + N / 2.0
        return (1 << N, f, (4.0 / 3.0) * 8.0 * ek3,
                "Schur-complement tree: 40 flops per updated entry of every included child node (torontonian.cu), "
                "averaged per subset; the kernel is shared-memory bound, the FP64 fraction is reported for scale only",
                "FP64 DFMA (vector pipe), operands in shared memory")
    if kind == "ltor":
        N = n // 2
        f = tor_tree_flops(N, aug=1) / float(1 << N)
        ek3 = N * (N - 1) * (N - 2) / 8.0 + 3 * N * (N - 1) / 4.0 + N / 2.0
        return (1 << N, f, (4.0 / 3.0) * 8.0 * ek3,
                "bordered Schur-complement tree (torontonian.cu, AUG = 1): 40 flops per updated entry of every included "
                "child node incl. the border row/column; shared-memory bound, the FP64 fraction is reported for scale only",
                "FP64 DFMA (vector pipe), operands in shared memory")
    if kind == "mtl":
        # reference: powertrace(Sigma_p, n + 1) = n - 1 products of the 2k x 2k sub-matrix per subset, k ~ Binomial(n, 1/2)
        ek3 = n * (n - 1) * (n - 2) / 8.0 + 3 * n * (n - 1) / 4.0 + n / 2.0
        ref = 8.0 * 8.0 * ek3 * (n - 1)
        return (1 << n, ref * ((n // 2) / max(1.0, n - 1.0)), ref,
                "reference: 8 (2k)^3 (n-1) per subset (product chain to tr Sigma_p^n), averaged over k ~ Bin(n, 1/2); the "
                "kernel pairs traces and needs ~n/2 products",
                "FP64 DFMA (vector pipe), warp per subset, operands in shared memory")
    if kind == "brs":
        per = (1 << (n - 1)) * (8.0 * n - 4) + 8.0 * n * n * (n / 2.0)
        return (1 << n, per, per,
                "per row subset Y: Gram matrix A_Y^H A_Y + E (8 n^2 |Y| flops, |Y| ~ n/2) and a Glynn permanent of it "
                "(2^(n-1) Gray steps x (8n - 4))",
                "FP64 DFMA (vector pipe), warp per row subset")
    raise ValueError(kind)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=200):
        self.rows, self.proc, self.index, self.period = [], None, index, int(period_ms)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", str(self.period)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# ---------------------------------------------------------------------------------------------------------------------------------
# flop models of the batched workloads
# ---------------------------------------------------------------------------------------------------------------------------------
def gbs_reference_flops(pats):
    """Reference-algorithm flops for the GBS patterns workload: a pattern with N_p photons is a loop hafnian of
    the reduction-expanded 2N_p x 2N_p matrix: 2^(N_p-1) Glynn subsets x 8 (2N_p)^3 (N_p - 1)."""
    Np = pats.sum(axis=1).astype(np.float64)
    Np = Np[Np >= 2]
    return float(np.sum(2.0 ** (Np - 1) * 8.0 * (2 * Np) ** 3 * (Np - 1)))


def gbs_executed_flops(rpt):
    """Flops the batched kernels EXECUTE for the repetition patterns ``rpt[B, nv]`` (even totals, loops): per pattern
    `steps` mixed-radix subsets of the UN-expanded pairing (matched_reps), each a pairing power-trace chain on the
    s = 2E matrix: floor((T-1)/2) products (8 s^3 each) for the traces up to T = N/2, plus the loop row's floor(T/2)
    row-times-matrix products (8 s^2) and the trace / pairing inner products (8 s^2 per needed trace).  Padding of the
    DMMA tiles is NOT counted (useful flops only)."""
    from thewalrus_b200._prep import glynn_steps, matched_reps

    uniq, counts = np.unique(np.asarray(rpt), axis=0, return_counts=True)
    total = 0.0
    for row, cnt in zip(uniq, counts):
        N = int(row.sum())
        if N < 2:
            continue
        _, er, odd = matched_reps([int(x) for x in row])
        s, T = 2 * len(er), N // 2
        steps = glynn_steps(er, True, odd is not None)
        nprod = (T - 1) // 2 if odd is None else max(T - 1, 0)
        total += cnt * steps * (8.0 * s**3 * nprod + 8.0 * s * s * (T // 2) + 8.0 * s * s * T)
    return total


def sampler_reference_flops(det, cutoff):
    """Reference-algorithm flops of the chains that produced the patterns ``det[S, M]``.  At mode i a chain needs the
    loop hafnians of the repetition patterns (n_1 .. n_(i-1), k), k = 0..cutoff.  Each is counted as the reference
    evaluates a repeated-vertex loop hafnian: pair the vertices (matched_reps), prod(edge_reps + 1) mixed-radix
    subsets (Glynn halves the first edge), per subset the product chain 8 s^3 (N/2 - 1) plus N/2 mat-vecs 8 s^2 on
    the s = 2 #edges reduced matrix (an upper bound: edges with delta = 0 drop out)."""
    from thewalrus_b200._prep import glynn_steps, matched_reps

    det = np.asarray(det, dtype=np.int64)
    S, M = det.shape
    total = 0.0
    for mode in range(M):
        pref, counts = np.unique(det[:, :mode], axis=0, return_counts=True) if mode else (np.zeros((1, 0), dtype=np.int64), np.array([S]))
        for row, cnt in zip(pref, counts):
            for k in range(cutoff + 1):
                rpt = list(row) + [k]
                N = int(sum(rpt))
                if N < 2:
                    continue
                _, er, odd = matched_reps(rpt)
                s_dim, T = 2 * len(er), N // 2
                steps = glynn_steps(er, True, odd is not None)
                total += cnt * steps * (8.0 * s_dim ** 3 * max(T - 1, 0) + 8.0 * s_dim ** 2 * T)
    return total


# ---------------------------------------------------------------------------------------------------------------------------------
# what both arms agree on: metric, units, config
# ---------------------------------------------------------------------------------------------------------------------------------
def metric_name(workload):
    if workload == "hafnian50":
        return METRIC
    if workload.startswith("gbs"):
        return f"{workload} GBS pattern probabilities/s"
    if workload.startswith("hsample"):
        return f"{workload} GBS photon-number samples/s"
    return f"{workload} subsets/s"


def gbs_inputs(workload, batch):
    import thewalrus_b200 as wb

    M = int(workload[3:])
    mu, cov, pats = make_gbs_state(M, batch, seed=1000 * 3 + M)
    A, gamma = wb.quantum._state(mu, cov, 2, 1e-10)
    rpt = np.ascontiguousarray(np.concatenate([pats, pats], axis=1))
    return M, mu, cov, pats, A, gamma, rpt


INPUT_TEXT = {
    "hafnian": "random complex symmetric G+G^T, seed 1000*config+n", "lhaf": "random complex symmetric G+G^T, loops = diagonal",
    "perm": "n x n block of a 2n Haar unitary",
    "tor": "O = I - Q^-1 of an N-mode GBS state (Haar interferometer, r=1.5, eta=0.8), all detectors click",
    "ltor": "O = I - sigma^-1, gamma = (sigma^-1 alpha)^* of the displaced N-mode GBS state, all detectors click",
    "mtl": "random complex symmetric 2n x 2n matrix / sqrt(8n)", "brs": "n x n block A of a 2n-mode Haar unitary, E = I - A^H A"}


def describe(workload, args):
    """(kind, n, units per step, unit, config dict) — identical for the GPU arm and the reference arm."""
    kind = workload.rstrip("0123456789")
    n = int(workload[len(kind):])
    world = max(1, args.gpus)
    if kind == "gbs":
        units, unit, what = args.batch, "patterns/s", "photon-number patterns"
        inp = (f"{n}-mode Gaussian state (Haar interferometer, r=0.5, eta=0.8, displaced), {units} Poisson(0.45) "
               "patterns with <= 10 photons")
        par = f"pattern shards x{world}, one all-gather"
        n = 2 * n
    elif kind == "hsample":
        units, unit, what = min(args.batch, 2048), "samples/s", "photon-number samples (accepted or not)"
        inp = (f"{n}-mode Gaussian state (Haar interferometer, r=0.4, eta=0.8, displaced), {units} chains per step "
               f"advanced together, cutoff {args.cutoff}")
        par = f"independent chains x{world} (replicas only, no collective)"
    else:
        units, unit, what = units_and_flops(kind, n)[0], "subsets/s", "subsets (reference `steps`)"
        inp = INPUT_TEXT[kind]
        par = f"subset-index shards x{world}, one all-reduce"
    cfg = {"workload": workload, "n": n, "units_per_step": units, "units": what, "input": inp, "parallelism": par,
           "l2_flush": True,
           "l2_note": "256 MiB fill between timed iterations; inputs are KB-sized, the path is FP64-pipe bound"}
    return kind, n, units, unit, cfg


# ---------------------------------------------------------------------------------------------------------------------------------
# CPU arms: the reference itself (numba, baseline/_ref) and the C port (oracle/)
# ---------------------------------------------------------------------------------------------------------------------------------
def host_threads():
    # all host cores this process may use — NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


_REF = {}


def load_reference():
    """Import the UNMODIFIED reference (baseline/_ref, installed by __graft_entry__.build()) with the dask stand-in.
    Returns (module, None) or (None, reason)."""
    if "mod" in _REF:
        return _REF["mod"], _REF.get("why")
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    why = None
    mod = None
    if not os.path.isdir(os.path.join(ref_dir, "thewalrus")):
        why = "baseline/_ref/thewalrus missing (run __graft_entry__.build() where /root/reference exists)"
    else:
        try:
            import numba  # noqa: F401

            for p in (os.path.join(ROOT, "baseline", "dask_shim"), ref_dir):
                if p not in sys.path:
                    sys.path.insert(0, p)
            import thewalrus as mod
        except Exception as exc:  # numba / llvmlite missing on the box, or the import fails
            mod, why = None, f"reference import failed: {type(exc).__name__}: {exc}"
    _REF["mod"], _REF["why"] = mod, why
    return mod, why


def _reference_range_harness():
    """the reference's own per-subset functions (find_kept_edges -> get_AX_S -> f, and the loop variants)
    in a numba prange with the scalar `H +=` reduction of _calc_hafnian (_hafnian.py:443-462), over a RANGE of subset
    indices so that n = 50 / 56 can be sampled; all reps 1, Glynn."""
    if "harness" in _REF:
        return _REF["harness"]
    import numba
    from thewalrus._hafnian import f, f_loop, find_kept_edges, get_AX_S, get_submatrices

    @numba.jit(nopython=True, parallel=True)
    def haf_range(A, edge_reps, j0, j1):
        N = 2 * edge_reps.sum()
        H = np.complex128(0)
        for j in numba.prange(j0, j1):
            kept = find_kept_edges(j, edge_reps)
            edge_sum = kept.sum()
            kept = 2 * kept - edge_reps
            AX_S = get_AX_S(kept, A)
            H += (-1.0) ** (N // 2 - edge_sum) * f(AX_S, N)[N // 2]
        return H

    @numba.jit(nopython=True, parallel=True)
    def lhaf_range(A, D, edge_reps, j0, j1):
        N = 2 * edge_reps.sum()
        H = np.complex128(0)
        for j in numba.prange(j0, j1):
            kept = find_kept_edges(j, edge_reps)
            edge_sum = kept.sum()
            kept = 2 * kept - edge_reps
            AX_S, XD_S, D_S, _ = get_submatrices(kept, A, D, np.zeros(len(D), dtype=np.complex128))
            H += (-1.0) ** (N // 2 - edge_sum) * f_loop(AX_S, XD_S, D_S, N)[N // 2]
        return H

    _REF["harness"] = (haf_range, lhaf_range)
    return _REF["harness"]


def cpu_reference(kind, n, X, seconds):
    """Time the reference's numba path on the host cores (bounded sample).  None if it does not apply / import."""
    tw, why = load_reference()
    if tw is None:
        return None, why
    import numba

    threads = numba.get_num_threads()
    if kind in ("hafnian", "lhaf"):
        loop = kind == "lhaf"
        total = 1 << (n // 2 - 1)
        if n <= 30:   # the whole call fits: thewalrus.hafnian itself
            tw.hafnian(X, loop=loop)                       # JIT / cache warm-up
            reps, t0 = 0, time.perf_counter()
            while True:
                tw.hafnian(X, loop=loop)
                reps += 1
                dt = time.perf_counter() - t0
                if dt >= seconds or reps >= 200:
                    break
            return {"value": reps * total / dt, "unit": "subsets/s", "cores": threads, "kind": "reference", "seconds": dt,
                    "sample": f"{reps} complete calls of the reference's thewalrus.hafnian(A{', loop=True' if loop else ''}) "
                              f"(numba prange over {threads} threads, OPENBLAS_NUM_THREADS=1)"}, None
        haf_range, lhaf_range = _reference_range_harness()
        from thewalrus._hafnian import matched_reps

        x, er, _ = matched_reps(np.ones(n, dtype=np.int64))
        Ax = np.ascontiguousarray(X[np.ix_(x, x)].astype(np.complex128))
        Dx = np.ascontiguousarray(np.diag(X)[x].astype(np.complex128))
        er = np.asarray(er, dtype=np.int64)
        run = (lambda a, b: lhaf_range(Ax, Dx, er, a, b)) if loop else (lambda a, b: haf_range(Ax, er, a, b))
        run(0, 4 * threads)                                 # JIT compile + thread pool
        t0 = time.perf_counter()
        run(0, 64 * threads)
        rate = 64 * threads / (time.perf_counter() - t0)
        sample = int(min(total, max(64 * threads, rate * seconds)))
        j0 = (total - sample) // 2                           # a window in the middle of the index space
        t0 = time.perf_counter()
        run(j0, j0 + sample)
        dt = time.perf_counter() - t0
        return {"value": sample / dt, "unit": "subsets/s", "cores": threads, "kind": "reference", "seconds": dt,
                "sample": f"{sample} of {total} subset indices from {j0} through the reference's own find_kept_edges -> "
                          f"get_AX_S -> f{'_loop' if loop else ''} (thewalrus/_hafnian.py) in a numba prange with its scalar H += "
                          f"reduction, {threads} threads, OPENBLAS_NUM_THREADS=1; extrapolated to the full sum"}, None
    if kind == "gbs":
        from thewalrus.quantum import density_matrix_element

        mu, cov, pats = X[3], X[4], X[5]
        density_matrix_element(mu, cov, list(pats[0]), list(pats[0]))    # JIT / cache warm-up
        order = np.random.default_rng(0).permutation(len(pats))          # a random sample of the batch: typical photon numbers
        done, t0 = 0, time.perf_counter()
        while done < len(order) and time.perf_counter() - t0 < seconds:
            p = [int(v) for v in pats[order[done]]]
            density_matrix_element(mu, cov, p, p)
            done += 1
        dt = time.perf_counter() - t0
        return {"value": done / dt, "unit": "patterns/s", "cores": threads, "kind": "reference", "seconds": dt,
                "sample": f"{done} randomly chosen patterns of the {len(pats)}, one reference density_matrix_element call each "
                          f"(quantum/fock_tensors.py:191-232; numba prange inside, {threads} threads)"}, None
    return None, f"no bounded-sample harness of the reference for '{kind}' (its loop has no range argument and the full call takes minutes); C port used"


def cpu_port(kind, n, X, seconds=15.0, cutoff=6):
    """Time the oracle's C port (reference algorithm) on the host cores over a bounded sample."""
    from oracle import c_oracle as co

    threads = host_threads()
    if kind in ("hafnian", "lhaf"):
        x = co.matched_order(X)
        Ax = np.ascontiguousarray(X[np.ix_(x, x)])
        Dx = np.ascontiguousarray(np.diag(X)[x]) if kind == "lhaf" else None
        total = 1 << (n // 2 - 1)
        co.hafnian_range(Ax, 0, min(total, 64 * threads), Dx, threads)  # warm-up (thread pool, page faults)
        t0 = time.perf_counter()
        co.hafnian_range(Ax, 0, min(total, 256 * threads), Dx, threads)
        rate = min(total, 256 * threads) / (time.perf_counter() - t0)
        sample = int(min(total, max(1024, rate * seconds)))
        t0 = time.perf_counter()
        co.hafnian_range(Ax, 0, sample, Dx, threads)
        dt = time.perf_counter() - t0
        what = f"first {sample} of {total} Glynn subsets of the same {n}x{n} matrix, full product-chain algorithm"
    elif kind == "perm":
        sample = int(min(1 << (n - 1), 4e7 * threads * seconds / 15.0))
        t0 = time.perf_counter()
        co.perm_range(X, 0, 0, sample, threads)
        dt = time.perf_counter() - t0
        what = f"first {sample} of {1 << (n - 1)} Gray-code steps (the reference itself is single-threaded; the port splits the range over threads)"
    elif kind == "tor" and n <= 48 and seconds >= 14.0:
        N = n // 2
        sample = 1 << N
        t0 = time.perf_counter()
        co.tor_recursive(X)
        dt = time.perf_counter() - t0
        threads = 1
        what = "full recursive torontonian (single thread, as the reference)"
    elif kind in ("tor", "ltor", "mtl", "brs"):   # tor: beyond 24 modes (full recursion > 10 min) or a short budget
        from oracle import walrus_oracle as wo

        total = 1 << (n // 2 if kind in ("tor", "ltor") else n)
        j0, sample, dt = total // 3, 0, 0.0       # a window in the middle of the index space: typical subset sizes
        step = {"tor": 100000, "ltor": 2000, "mtl": 2000, "brs": 8}[kind]
        t0 = time.perf_counter()
        used = 1
        if kind == "ltor":
            step, used = 20000 * threads, threads     # numba_ltor is a parallel prange in the reference: all cores
        elif kind == "brs":
            step = 64                                   # C port, one thread: the reference's brs loop is serial
        while dt < seconds and j0 + sample < total:
            a, b = j0 + sample, min(total, j0 + sample + step)
            if kind == "tor":
                co.tor_direct(X, a, b, 1)
            elif kind == "ltor":
                co.ltor_direct(X[0], X[1], a, b, threads)
            elif kind == "mtl":
                wo.montrealer(np.vstack([X[n:], X[:n]]), None, a, b)     # Xmat(n) @ A, as mtl does
            else:
                co.brs(X[0], X[1], a, b, 1)
            sample += b - a
            dt = time.perf_counter() - t0
        threads = used
        what = (f"{sample} of {total} subsets starting at index {j0} through the "
                + ("C + OpenMP port of numba_ltor (all host cores, as the reference's prange)" if kind == "ltor" else
                   "C port of brs (one thread, as the reference)" if kind == "brs" else
                   "C port of numba_tor (one thread, as the reference)" if kind == "tor" else
                   "NumPy restatement of the reference (single thread, as the reference)"))
    elif kind == "hsample":
        # the reference's chain (one sample at a time, one loop_hafnian_batch per mode) with the oracle as its kernel
        from thewalrus_b200 import quantum as wq
        from thewalrus_b200 import samples as wsamples

        def oracle_patterns(A, gamma, rpt, glynn=True, *, gamma_index=None, A_index=None, group=None, device=None):
            gamma = np.atleast_2d(gamma)        # one chain at a time: a single gamma row
            assert len(gamma) == 1
            return co.lhaf_patterns(A, gamma[0], rpt, glynn, threads)

        saved, wq.lhaf_patterns = wq.lhaf_patterns, oracle_patterns
        try:
            ch = wsamples._Chain(X[1], X[0], 2)
            t0 = time.perf_counter()
            sample = 0
            while time.perf_counter() - t0 < seconds:
                wsamples._hafnian_chains(ch, 1, cutoff, None)
                sample += 1
            dt = time.perf_counter() - t0
        finally:
            wq.lhaf_patterns = saved
        return {"value": sample / dt, "unit": "samples/s", "cores": threads, "kind": "port", "seconds": dt,
                "sample": f"{sample} chains, one at a time as the reference walks them, the cutoff + 1 loop hafnians of a "
                          "mode step through the C port (OpenMP over the outcomes)"}
    else:  # gbs: X = (A, gamma, rpt, ...); the C port of the reference's per-pattern loop hafnian, patterns over all cores
        A, gamma, rpt = X[0], X[1], X[2]
        block, sample = 256 * threads, 0
        co.lhaf_patterns(A, gamma, rpt[:threads], True, threads)       # warm-up (thread pool)
        t0 = time.perf_counter()
        while sample < len(rpt) and time.perf_counter() - t0 < seconds:
            co.lhaf_patterns(A, gamma, rpt[sample:sample + block], True, threads)
            sample = min(len(rpt), sample + block)
        dt = time.perf_counter() - t0
        what = (f"first {sample} of {len(rpt)} patterns, one loop hafnian per pattern through the C port of the reference "
                "algorithm, patterns spread over all host cores (the reference makes one Python call per pattern, numba "
                "prange inside)")
        return {"value": sample / dt, "unit": "patterns/s", "cores": threads, "kind": "port", "sample": what, "seconds": dt}
    return {"value": sample / dt, "unit": "subsets/s", "cores": threads, "kind": "port", "sample": what,
            "seconds": dt}


def workload_inputs(workload, args):
    kind = workload.rstrip("0123456789")
    if kind == "gbs":
        M, mu, cov, pats, A, gamma, rpt = gbs_inputs(workload, args.batch)
        return "gbs", 2 * M, (A, gamma, rpt, mu, cov, pats)
    return make_input(workload)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import build as obuild

    obuild.ensure()
    kind, n, units, unit, cfg = describe(args.workload, args)
    _, _, X = workload_inputs(args.workload, args)
    per = args.seconds if args.seconds > 0 else max(2.0, 60.0 / max(1, args.warmup + args.steps))
    vals, ports = [], []
    why = None
    for i in range(args.warmup + args.steps):
        cb, why = (None, "forced: --cpu-kind port") if args.cpu_kind == "port" else cpu_reference(kind, n, X, per)
        if cb is None:
            cb = cpu_port(kind, n, X, seconds=per, cutoff=args.cutoff)
        if i >= args.warmup:
            vals.append(cb)
    if args.cpu_kind == "both" and vals[-1]["kind"] == "reference":
        ports.append(cpu_port(kind, n, X, seconds=per, cutoff=args.cutoff))
    v = statistics.mean(c["value"] for c in vals)
    cfg = dict(cfg)
    cfg["note"] = "reference arm: each step is a bounded sample of the workload; ms_per_step is extrapolated to the full step"
    cbase = {"value": v, "unit": unit, "cores": vals[-1]["cores"], "kind": vals[-1]["kind"], "sample": vals[-1]["sample"],
             "host_cpus": host_threads()}
    if why:
        cbase["reference_unavailable"] = why
    if ports:
        cbase["port"] = ports[-1]
    line = {"impl": "reference", "metric": metric_name(args.workload),
            "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": units / v * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg, "cpu_baseline": cbase,
            "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess(workload, args, seconds):
    """Run the reference arm of this file in a fresh process (numba threading and OPENBLAS_NUM_THREADS=1 must be set
    before numpy loads; torchrun's OMP_NUM_THREADS=1 must not leak into the OpenMP port) and return its cpu_baseline."""
    env = {k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "RANK", "LOCAL_RANK", "WORLD_SIZE")}
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload, "--steps", "1",
           "--warmup", "0", "--seconds", str(seconds), "--batch", str(args.batch), "--cutoff", str(args.cutoff),
           "--cpu-kind", "both"]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
        for ln in out.stdout.splitlines():
            if ln.startswith("{"):
                return json.loads(ln)["cpu_baseline"]
        return {"error": (out.stderr or out.stdout)[-400:]}
    except Exception as exc:  # noqa: BLE001
        return {"error": f"{type(exc).__name__}: {exc}"}


# ---------------------------------------------------------------------------------------------------------------------------------
# goldens: the error of the complete result, printed in the line
# ---------------------------------------------------------------------------------------------------------------------------------
def golden_error(workload, res, extra=None):
    """Relative error of the full result of this run against tests/golden/reference_fullsize.json (None if the golden
    has no entry for the workload)."""
    try:
        with open(GOLDEN) as fh:
            g = json.load(fh)
    except OSError:
        return None

    def cz(d):
        return complex(d["re"], d["im"]) if isinstance(d, dict) else complex(d)

    def rel(a, b):
        return abs(a - b) / max(abs(b), 1e-300)

    out = {}
    if workload in ("hafnian50", "hafnian24") and workload in g:
        e = g[workload]
        for key, name in (("oracle_ld", "long-double C oracle"), ("oracle_double", "double C oracle (full product chain, Kahan)"),
                          ("reference", "reference numba thewalrus.hafnian")):
            if key in e:
                out[name] = rel(res, cz(e[key]))
    elif workload == "perm32" and "perm32" in g:
        e = g["perm32"]
        out["long-double C oracle (2^31 steps)"] = rel(res, cz(e["oracle_ld"]))
        if "reference_bbfg" in e:
            out["reference perm(bbfg)"] = rel(res, cz(e["reference_bbfg"]))
    elif workload == "tor48" and "tor48" in g:
        e = g["tor48"]
        out["long-double C oracle"] = rel(res.real, e["oracle_ld"])
        out["reference rec_torontonian"] = rel(res.real, cz(e["reference_rec"]).real)
    elif workload == "gbs16" and "gbs16" in g and extra is not None and len(extra) == g["gbs16"]["B"]:
        npy = os.path.join(os.path.dirname(GOLDEN), "gbs16_probabilities_ld.npy")
        if os.path.exists(npy):
            want = np.maximum(np.load(npy), 0.0)       # probabilities() clips at 0 (fock_tensors.py:424-428)
            out["sum of all probabilities vs long-double C oracle"] = rel(float(np.sum(np.sort(extra))), float(np.sum(np.sort(want))))
            out["worst single probability (scaled by max(p, 1e-3 max p)) vs long-double C oracle"] = float(
                np.max(np.abs(extra - want) / np.maximum(want, 1e-3 * want.max())))
    return out or None


# ---------------------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------------------
def gpu_arm(workload, args, env, steps, warmup, min_seconds=0.0):
    """Measure one workload on the GPU(s).  ``min_seconds``: repeat steps until the timed region has run at least this
    long (short workloads: the nvidia-smi sampler needs time to see the clocks)."""
    torch, dist, lib, dev = env["torch"], env["dist"], env["lib"], env["dev"]
    world, rank, local = env["world"], env["rank"], env["local"]
    import thewalrus_b200 as wb
    from thewalrus_b200 import _engine, _lib
    from thewalrus_b200._prep import shard_range

    kind, n, units, unit, cfg = describe(workload, args)
    is_gbs = kind == "gbs"
    if is_gbs:
        _, _, X = workload_inputs(workload, args)
        A, gamma, rpt, mu, cov, pats = X
        ref_flops = gbs_reference_flops(pats) / units
        my_flops = gbs_executed_flops(rpt) / units
        model = ("EXECUTED useful flops of the batched kernels (bench.gbs_executed_flops): per pattern prod(edge_reps+1) "
                 "mixed-radix subsets of the un-expanded pairing, each floor((T-1)/2) products 8 s^3 + loop row + pairing "
                 "inner products on the s = 2E matrix; DMMA tile padding not counted")
        pipe = "FP64 DMMA.8x8x4 (tensor pipe) for even patterns; FP64 DFMA warp-per-subset kernel for odd ones"
        lo, hi = shard_range(units, rank, world)
    elif kind == "hsample":
        _, _, X = make_input(workload)
        lo, hi = shard_range(units, rank, world)
        from thewalrus_b200 import samples as wsamples

        chain = wsamples._Chain(X[1], X[0], 2)
        my_flops = ref_flops = 0.0             # filled from the patterns actually drawn (below)
        model = ("reference-algorithm flops of the loop hafnians a chain evaluates: per mode and outcome k one repeated-"
                 "vertex loop hafnian = prod(edge_reps + 1) mixed-radix subsets x (8 s^3 (N/2 - 1) + 8 s^2 N/2) on the "
                 "s = 2 #edges reduced matrix; these are small problems, the step is launch/latency bound")
        pipe = "FP64 DFMA (vector pipe), warp per subset, operands in shared memory"
    else:
        _, _, X = make_input(workload)
        _, my_flops, ref_flops, model, pipe = units_and_flops(kind, n)
        lo, hi = shard_range(units if kind not in ("tor", "ltor") else _engine.tor_num_prefixes(n // 2), rank, world)

    # ---- device-resident inputs for the kernel-only number
    dA = dD = None
    wsb = 8
    if kind in ("hafnian", "lhaf"):
        x, er, _ = wb.matched_reps([1] * n)
        Ax = np.ascontiguousarray(X[np.ix_(x, x)].astype(np.complex128))
        dA = torch.from_numpy(Ax.view(np.float64).reshape(-1)).to(dev)
        dD = torch.from_numpy(np.ascontiguousarray(np.diag(X)[x]).view(np.float64).reshape(-1)).to(dev) if kind == "lhaf" else None
        wsb = lib.wb200_hafnian_workspace_bytes(n)
        launches_per_step = 2      # haf_dmma (builds its fragment table itself), final_reduce
    elif kind == "perm":
        dA = torch.from_numpy(np.ascontiguousarray(X).view(np.float64).reshape(-1)).to(dev)
        wsb = lib.wb200_perm_workspace_bytes(n)
        launches_per_step = 2      # perm_kernel, final_reduce
    elif kind == "tor":
        dA = torch.from_numpy(np.ascontiguousarray(X, dtype=np.complex128).view(np.float64).reshape(-1)).to(dev)
        wsb = lib.wb200_tor_workspace_bytes(n // 2)
        launches_per_step = 3      # tor_prep, tor_kernel, final_reduce
    elif kind == "ltor":
        dA = torch.from_numpy(np.ascontiguousarray(X[0], dtype=np.complex128).view(np.float64).reshape(-1)).to(dev)
        dD = torch.from_numpy(np.ascontiguousarray(X[1], dtype=np.complex128).view(np.float64).reshape(-1)).to(dev)
        wsb = lib.wb200_tor_workspace_bytes(n // 2)
        launches_per_step = 3      # tor_prep, tor_kernel<1>, final_reduce
    elif kind == "mtl":
        launches_per_step = 2      # mtl_kernel, final reduction
    elif kind == "brs":
        launches_per_step = 2      # brs_kernel, final reduction
    elif kind == "hsample":
        launches_per_step = 9 * n  # per mode: shift, pattern build, prep, scan, ~3 class kernels, final, draw
    else:
        launches_per_step = 7      # pat_prep, scan, pat_dmma x4 (tile classes of E <= 10), pat_final
    host_entry = kind in ("gbs", "mtl", "brs", "hsample")   # *_host entry points: they time their own launches
    drawn = []
    ws = torch.empty((wsb + 7) // 8, dtype=torch.float64, device=dev)
    out = torch.zeros(4, dtype=torch.float64, device=dev)
    table = torch.zeros((world, 4), dtype=torch.float64, device=dev)
    flush = env["flush"]
    stream = torch.cuda.current_stream(dev)
    inner_ms = []   # kernel time reported by the *_host entry points (CUDA events around the launches only)
    gbs_out = [None]

    def kernel_step():
        if kind in ("hafnian", "lhaf"):
            rc = lib.wb200_hafnian_dev(dA.data_ptr(), dD.data_ptr() if dD is not None else None, n, lo, hi, out.data_ptr(),
                                       ws.data_ptr(), ws.numel() * 8, stream.cuda_stream)
        elif kind == "perm":
            rc = lib.wb200_perm_dev(dA.data_ptr(), n, 0, lo, hi, out.data_ptr(), ws.data_ptr(), ws.numel() * 8,
                                    stream.cuda_stream)
        elif kind == "tor":
            rc = lib.wb200_tor_dev(dA.data_ptr(), n // 2, lo, hi, out.data_ptr(), ws.data_ptr(), ws.numel() * 8,
                                   stream.cuda_stream)
        elif kind == "ltor":
            rc = lib.wb200_ltor_dev(dA.data_ptr(), dD.data_ptr(), n // 2, lo, hi, out.data_ptr(), ws.data_ptr(),
                                    ws.numel() * 8, stream.cuda_stream)
        elif kind in ("mtl", "brs", "hsample"):
            _engine.kernel_ms_log = []
            if kind == "mtl":
                _engine.mtl_range(X, None, lo, hi, dev)
            elif kind == "brs":
                _engine.brs_range(X[0], X[1], lo, hi, dev)
            else:
                drawn.append(wsamples._hafnian_chains(chain, hi - lo, args.cutoff, dev))
            inner_ms.append(sum(_engine.kernel_ms_log))
            _engine.kernel_ms_log = None
            rc = 0
        else:
            gbs_out[0], ms = _engine.lhaf_patterns_local(A, gamma, rpt[lo:hi], True, dev, want_ms=True)
            inner_ms.append(ms)
            rc = 0
        _lib.check(rc, "kernel step")
        if world > 1 and not host_entry:  # the one collective of the path: all-reduce of the (hi, lo) partials
            table.zero_()
            table[rank] = out
            dist.all_reduce(table)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(warmup):
        kernel_step()
    sync_all()
    if min_seconds > 0:       # size the timed region from one timed warm step
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        kernel_step()
        e1.record(stream)
        sync_all()
        one = max(inner_ms[-1] if inner_ms else e0.elapsed_time(e1), 0.02) * 1e-3
        steps = int(min(4000, max(steps, math.ceil(min_seconds / one))))
    sampler = ClockSampler(local, period_ms=50 if min_seconds > 0 else 200)
    if rank == 0:
        sampler.start()
    total_ms = 0.0
    kern_ms = []
    for _ in range(steps):
        flush.fill_(1.0)  # L2 flush between timed iterations (inputs are KB-sized, far below L2)
        sync_all()
        inner_ms.clear()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        kernel_step()
        e1.record(stream)
        sync_all()
        ms = inner_ms[-1] if inner_ms else e0.elapsed_time(e1)   # host-buffer entry points time their own launches
        kern_ms.append(ms)
        total_ms += ms
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = units * steps / (total_ms * 1e-3)
    if host_entry:
        res = 0j
    elif world > 1:
        res = _engine.combine4(table.cpu().numpy())
    else:
        res = _engine.combine4([out.cpu().numpy()])
    if kind in ("hafnian", "lhaf"):
        res_scaled = res * 0.5 ** (n // 2 - 1)
    elif kind == "perm":
        res_scaled = res / float(1 << (n - 1))
    else:
        res_scaled = res

    # ---- end to end through the public API with host buffers
    grp = True if world > 1 else None

    def e2e_step():
        if kind == "hafnian":
            return wb.hafnian(X, group=grp)
        if kind == "lhaf":
            return wb.hafnian(X, loop=True, group=grp)
        if kind == "perm":
            return wb.perm(X, method="glynn", group=grp)
        if kind == "tor":
            return wb.tor(X, group=grp)
        if kind == "ltor":
            return wb.ltor(X[0], X[1], group=grp)
        if kind == "mtl":
            return wb.mtl(X, group=grp)
        if kind == "brs":
            return wb.brs(X[0], X[1], group=grp)
        if kind == "hsample":   # every rank draws its share of the samples (independent chains, no collective)
            got = wsamples.hafnian_sample_state(X[1], hi - lo, mean=X[0], cutoff=args.cutoff, max_photons=10 ** 6)
            return float(got.sum())
        return wb.probabilities_batch(mu, cov, pats, group=grp)

    r_e2e = e2e_step()
    sync_all()
    step_s = total_ms * 1e-3 / steps
    if min_seconds > 0:
        e2e_steps = int(min(4000, max(3, math.ceil(0.5 * min_seconds / max(1e-5, step_s)))))
    else:            # multi-second steps: a handful of end-to-end calls is enough (the device-timed K steps stay exact)
        e2e_steps = steps if step_s < 0.5 else min(steps, 5)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        r_e2e = e2e_step()
    sync_all()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        return None

    peak = ctypes.c_double(0)
    dmma = kind in ("hafnian", "lhaf", "gbs")
    lib.wb200_fp64_peak(local, 1 if dmma else 0, ctypes.byref(peak))
    per_gpu_units = (hi - lo) if kind not in ("tor", "ltor") else units / world
    if kind == "hsample":
        my_flops = ref_flops = sampler_reference_flops(np.concatenate(drawn), args.cutoff) / max(1, len(drawn) * (hi - lo))
    kms = statistics.mean(kern_ms)
    achieved = per_gpu_units * my_flops / (kms * 1e-3) * 1e-12
    if is_gbs:
        h2d, d2h = int(mu.nbytes + cov.nbytes + pats.nbytes), int(8 * len(pats))
        api = "thewalrus_b200.probabilities_batch(mu, cov, patterns) with host NumPy arrays"
    elif kind == "hsample":
        K = args.cutoff + 1  # per mode step: B block + one gamma row per chain + patterns in, lhafs out
        h2d = int(sum(16 * m * m + (hi - lo) * (16 * m + 4 * m * K + 4 * K) for m in range(1, n + 1)))
        d2h = int(n * (hi - lo) * K * 16)
        api = "thewalrus_b200.samples.hafnian_sample_state(cov, S, mean=mu, cutoff) with host NumPy arrays"
    else:
        Xs = X if isinstance(X, tuple) else (X,)
        h2d, d2h = int(sum(np.asarray(x).nbytes for x in Xs)), 32
        api = {"hafnian": "thewalrus_b200.hafnian(A)", "lhaf": "thewalrus_b200.hafnian(A, loop=True)",
               "perm": "thewalrus_b200.perm(A, method='glynn')", "tor": "thewalrus_b200.tor(O)",
               "ltor": "thewalrus_b200.ltor(O, gamma)", "mtl": "thewalrus_b200.mtl(A)",
               "brs": "thewalrus_b200.brs(A, E)"}[kind] + " with a host NumPy array"
    traffic = MEASURED_TRAFFIC.get(workload)
    if is_gbs:
        full_res, extra = None, np.asarray(r_e2e)
        r_e2e_val = float(np.sum(extra))
    else:
        full_res, extra, r_e2e_val = res_scaled, None, complex(r_e2e)
    line = {
        "metric": metric_name(workload),
        "value": value, "unit": unit, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "e2e": {"value": units * e2e_steps / e2e_s, "unit": unit, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / e2e_steps * 1e3, "steps": e2e_steps, "api": api},
        "gpu_launches": launches_per_step * steps * world,
        "clocks": clocks,
        "roofline": {"bound": "tensor" if dmma else "fp64", "pipe": pipe,
                     "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s", "frac": achieved / peak.value,
                     "peak_source": "measured live by wb200_fp64_peak (dependent-chain micro-benchmark on the same pipe: "
                                    "DMMA for the hafnian kernels, DFMA otherwise); MEASURED_PEAKS.json has no FP64 entry; "
                                    "nominal 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2",
                     "flops_per_unit": my_flops, "flops_model": model,
                     "reference_algorithm_equivalent_tflops": per_gpu_units * ref_flops / (kms * 1e-3) * 1e-12,
                     "kernel_ms": kms, "traffic": traffic[0] if traffic else None,
                     "traffic_source": traffic[1] if traffic else None,
                     "traffic_note": "bytes per launch of the dominant kernel, dram__bytes_read.sum + dram__bytes_write.sum from "
                                     "the named ncu capture; the algorithmic traffic is KB (matrix in, 4 doubles per CTA out) — these "
                                     "paths are FP64-pipe bound"},
        "result": ({"sum_of_probabilities": r_e2e_val} if is_gbs else
                   {"re": complex(full_res).real, "im": complex(full_res).imag, "e2e_re": r_e2e_val.real, "e2e_im": r_e2e_val.imag}),
    }
    if not host_entry or is_gbs:
        err = golden_error(workload, full_res, extra)
        if err is not None:
            line["result_rel_err"] = max(err.values()) if is_gbs else min(err.values())
            line["result_rel_err_vs"] = err
            line["result_golden"] = "tests/golden/reference_fullsize.json (tests/golden/make_golden_fullsize.py)"
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="hafnian50")
    ap.add_argument("--batch", type=int, default=100000, help="patterns per step of the gbs workload")
    ap.add_argument("--cutoff", type=int, default=6, help="per-mode photon cutoff of the hsample workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other BASELINE configs in the default run")
    ap.add_argument("--seconds", type=float, default=0.0, help="reference arm: CPU seconds per step (default 60 / steps)")
    ap.add_argument("--cpu-kind", default="auto", choices=["auto", "port", "both"],
                    help="reference arm: numba reference when available (auto), the C port, or both")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    from thewalrus_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    env = {"torch": torch, "dist": dist, "lib": _lib.load(), "dev": dev, "world": world, "rank": rank, "local": local,
           "flush": torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)}  # > 126 MB L2
    args.gpus = world
    line = gpu_arm(args.workload, args, env, args.steps, args.warmup)
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_subprocess(args.workload, args, 15.0)
        if world == 1 and args.workload == "hafnian50" and not args.no_secondary:
            sec = {}
            for w in SECONDARY:
                try:
                    s = gpu_arm(w, args, env, 5, 3, min_seconds=2.0)
                    if not args.no_cpu_baseline:
                        s["cpu_baseline"] = cpu_baseline_subprocess(w, args, 5.0)
                    sec[w] = s
                except Exception as exc:  # noqa: BLE001  (a secondary config must never cost the headline line)
                    sec[w] = {"error": f"{type(exc).__name__}: {exc}"}
            line["secondary"] = sec
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
