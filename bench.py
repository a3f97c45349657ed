#!/usr/bin/env python
"""Headline benchmark: BASELINE.json metric "hafnian n=50 complex128 subsets/s & wall-time at 1/2/4/8 B200;
% FP64 peak".

One "step" = one complete hafnian of a random 50x50 complex128 symmetric matrix (2^24 Glynn subsets,
seed 1000*1+50 as in SURVEY.md 8d).  With N GPUs the subset index is sharded in contiguous ranges (strong
scaling: total work fixed) and the partial sums are combined by one all-reduce inside the timed region.

  python bench.py --gpus N --steps K --warmup W                 # GPU arm (torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K --warmup W  # CPU arm: the oracle port on host cores

Other workloads (the remaining BASELINE configs and the north-star targets):
--workload hafnian24|hafnian56|lhaf50|perm32|perm40|tor48|gbs16
and the SURVEY 8(f) components: ltor48 (loop torontonian), mtl14 (montrealer, 14 modes), brs12 (Bristolian of a
12 x 12 block), hsample8 (batched chain-rule photon-number sampler, 8 modes; unit: samples/s).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hafnian n=50 complex128 subsets/s"
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel (ncu captures in profiles/)
MEASURED_TRAFFIC = {"hafnian50": 5864960 + 59578624, "perm32": 79104, "tor48": 79104}


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


def units_and_flops(kind, n):
    """(units per step, executed-algorithm flops per unit of THIS implementation, reference-algorithm flops per
    unit, text of the model, FP64 pipe the kernel issues to)."""
    if kind in ("hafnian", "lhaf"):
        m = n // 2
        nprod = (m - 1) // 2
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

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
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


def gbs_reference_flops(pats):
    """Reference-algorithm flops for the GBS patterns workload: a pattern with N_p photons is a loop hafnian of
    the reduction-expanded 2N_p x 2N_p matrix: 2^(N_p-1) Glynn subsets x 8 (2N_p)^3 (N_p - 1)."""
    Np = pats.sum(axis=1).astype(np.float64)
    Np = Np[Np >= 2]
    return float(np.sum(2.0 ** (Np - 1) * 8.0 * (2 * Np) ** 3 * (Np - 1)))


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


def cpu_baseline(kind, n, X, seconds=15.0, cutoff=6):
    """Time the oracle's C port (reference algorithm) on the host cores over a bounded sample."""
    from oracle import c_oracle as co

    # all host cores this process may use — NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or co.max_threads())
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
    elif kind == "tor" and n <= 48:
        N = n // 2
        sample = 1 << N
        t0 = time.perf_counter()
        co.tor_recursive(X)
        dt = time.perf_counter() - t0
        threads = 1
        what = "full recursive torontonian (single thread, as the reference)"
    elif kind in ("tor", "ltor", "mtl", "brs"):   # tor: only beyond 24 modes, where the full recursion takes > 10 min
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
    else:  # gbs: X = (A, gamma, rpt); the C port of the reference's per-pattern loop hafnian, patterns over all cores
        A, gamma, rpt = X
        block, sample = 256 * threads, 0
        co.lhaf_patterns(A, gamma, rpt[:threads], True, threads)       # warm-up (thread pool)
        t0 = time.perf_counter()
        while sample < len(rpt) and time.perf_counter() - t0 < seconds:
            co.lhaf_patterns(A, gamma, rpt[sample:sample + block], True, threads)
            sample = min(len(rpt), sample + block)
        dt = time.perf_counter() - t0
        what = (f"first {sample} of {len(rpt)} patterns, one loop hafnian per pattern through the C port of the reference "
                "algorithm, patterns spread over all host cores (the reference makes one Python call per pattern, numba "
                "prange inside; it measured 177 patterns/s on 8 cores)")
        return {"value": sample / dt, "unit": "patterns/s", "cores": threads, "kind": "port", "sample": what, "seconds": dt}
    return {"value": sample / dt, "unit": "subsets/s", "cores": threads, "kind": "port", "sample": what,
            "seconds": dt}


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


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as ge  # noqa: F401  (builds the oracle if needed; CPU only)

    from oracle import build as obuild

    obuild.ensure()
    per = max(2.0, 60.0 / max(1, args.warmup + args.steps))
    if args.workload.startswith("gbs"):
        M, mu, cov, pats, A, gamma, rpt = gbs_inputs(args.workload, args.batch)
        kind, n, X, units, ref_flops, unit = "gbs", 2 * M, (A, gamma, rpt), len(rpt), gbs_reference_flops(pats) / len(rpt), "patterns/s"
    elif args.workload.startswith("hsample"):
        kind, n, X = make_input(args.workload)
        units, ref_flops, unit = min(args.batch, 2048), 0.0, "samples/s"
    else:
        kind, n, X = make_input(args.workload)
        units, _, ref_flops, _, _ = units_and_flops(kind, n)
        unit = "subsets/s"
    vals = []
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(kind, n, X, seconds=per, cutoff=args.cutoff)
        if i >= args.warmup:
            vals.append(cb)
    v = statistics.mean(c["value"] for c in vals)
    line = {"impl": "reference", "metric": metric_name(args.workload),
            "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": units / v * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "n": n, "units_per_step": units,
                       "note": "each step is a bounded sample of the workload; ms_per_step is extrapolated to the full step"},
            "cpu_baseline": {"value": v, "unit": unit, "cores": vals[-1]["cores"], "kind": "port",
                             "sample": vals[-1]["sample"], "gflops_reference_algorithm": v * ref_flops * 1e-9},
            "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


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
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    import thewalrus_b200 as wb
    from thewalrus_b200 import _engine, _lib
    from thewalrus_b200._prep import shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    is_gbs = args.workload.startswith("gbs")
    if is_gbs:
        M, mu, cov, pats, A, gamma, rpt = gbs_inputs(args.workload, args.batch)
        kind, n, X = "gbs", 2 * M, (A, gamma, rpt)
        units = len(rpt)
        ref_flops = gbs_reference_flops(pats) / units
        my_flops = ref_flops
        model = ("reference-algorithm flops of the reduction-expanded Glynn loop hafnians, sum_p 2^(N_p-1) 8 (2N_p)^3 (N_p-1); "
                 "the kernel works on the un-expanded matrices with mixed-radix subsets, so this is an equivalent, not an executed count")
        pipe = "FP64 DFMA (vector pipe), warp per subset, operands in shared memory"
        unit = "patterns/s"
        lo, hi = shard_range(units, rank, world)
    elif args.workload.startswith("hsample"):
        kind, n, X = make_input(args.workload)
        units = min(args.batch, 2048)          # chains per step, split over the ranks
        unit = "samples/s"
        lo, hi = shard_range(units, rank, world)
        from thewalrus_b200 import samples as wsamples

        chain = wsamples._Chain(X[1], X[0], 2)
        my_flops = ref_flops = 0.0             # filled from the patterns actually drawn (below)
        model = ("reference-algorithm flops of the loop hafnians a chain evaluates: per mode and outcome k one repeated-"
                 "vertex loop hafnian = prod(edge_reps + 1) mixed-radix subsets x (8 s^3 (N/2 - 1) + 8 s^2 N/2) on the "
                 "s = 2 #edges reduced matrix; these are small problems, the step is launch/latency bound")
        pipe = "FP64 DFMA (vector pipe), warp per subset, operands in shared memory"
    else:
        kind, n, X = make_input(args.workload)
        units, my_flops, ref_flops, model, pipe = units_and_flops(kind, n)
        unit = "subsets/s"
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
        launches_per_step = 3      # haf_prep, haf_dmma, final_reduce
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
        launches_per_step = 4 * n  # one patterns call (prep, scan, main, final) per mode
    else:
        launches_per_step = 4      # pat_prep, scan, pat_main, pat_final
    host_entry = kind in ("gbs", "mtl", "brs", "hsample")   # *_host entry points: they time their own launches
    drawn = []
    ws = torch.empty((wsb + 7) // 8, dtype=torch.float64, device=dev)
    out = torch.zeros(4, dtype=torch.float64, device=dev)
    table = torch.zeros((world, 4), dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream(dev)
    inner_ms = []   # kernel time reported by the *_host entry points (CUDA events around the launches only)

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
            _, ms = _engine.lhaf_patterns_local(A, gamma, rpt[lo:hi], True, dev, want_ms=True)
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

    for _ in range(args.warmup):
        kernel_step()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms = 0.0
    kern_ms = []
    for _ in range(args.steps):
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
    value = units * args.steps / (total_ms * 1e-3)
    if host_entry:
        res = 0j
    elif world > 1:
        res = _engine.combine4(table.cpu().numpy())
    else:
        res = _engine.combine4([out.cpu().numpy()])

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
        return float(np.sum(wb.probabilities_batch(mu, cov, pats, group=grp)))

    e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r_e2e = e2e_step()
    sync_all()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peak = ctypes.c_double(0)
        lib.wb200_fp64_peak(local, 1 if kind in ("hafnian", "lhaf") else 0, ctypes.byref(peak))
        per_gpu_units = (hi - lo) if kind not in ("tor", "ltor") else units / world
        if kind == "hsample":
            my_flops = ref_flops = sampler_reference_flops(np.concatenate(drawn), args.cutoff) / max(1, len(drawn) * (hi - lo))
        kms = statistics.mean(kern_ms)
        achieved = per_gpu_units * my_flops / (kms * 1e-3) * 1e-12
        if is_gbs:
            h2d, d2h = int(mu.nbytes + cov.nbytes + pats.nbytes), int(8 * len(pats))
            cfg_in = (f"{M}-mode Gaussian state (Haar interferometer, r=0.5, eta=0.8, displaced), {units} Poisson(0.45) "
                      "patterns with <= 10 photons")
            api = "thewalrus_b200.probabilities_batch(mu, cov, patterns) with host NumPy arrays"
        elif kind == "hsample":
            # per mode step: B block + one gamma row per chain + patterns in, lhafs out
            K = args.cutoff + 1
            h2d = int(sum(16 * m * m + (hi - lo) * (16 * m + 4 * m * K + 4 * K) for m in range(1, n + 1)))
            d2h = int(n * (hi - lo) * K * 16)
            cfg_in = (f"{n}-mode Gaussian state (Haar interferometer, r=0.4, eta=0.8, displaced), {units} chains per step "
                      f"advanced together, cutoff {args.cutoff}")
            api = "thewalrus_b200.samples.hafnian_sample_state(cov, S, mean=mu, cutoff) with host NumPy arrays"
        else:
            Xs = X if isinstance(X, tuple) else (X,)
            h2d, d2h = int(sum(np.asarray(x).nbytes for x in Xs)), 32
            cfg_in = {"hafnian": "random complex symmetric G+G^T, seed 1000*config+n", "lhaf": "random complex symmetric G+G^T, loops = diagonal",
                      "perm": "n x n block of a 2n Haar unitary", "tor": "O = I - Q^-1 of an N-mode GBS state (Haar interferometer, r=1.5, eta=0.8), all detectors click",
                      "ltor": "O = I - sigma^-1, gamma = (sigma^-1 alpha)^* of the displaced N-mode GBS state, all detectors click",
                      "mtl": "random complex symmetric 2n x 2n matrix / sqrt(8n)",
                      "brs": "n x n block A of a 2n-mode Haar unitary, E = I - A^H A"}[kind]
            api = {"hafnian": "thewalrus_b200.hafnian(A)", "lhaf": "thewalrus_b200.hafnian(A, loop=True)",
                   "perm": "thewalrus_b200.perm(A, method='glynn')", "tor": "thewalrus_b200.tor(O)",
                   "ltor": "thewalrus_b200.ltor(O, gamma)", "mtl": "thewalrus_b200.mtl(A)",
                   "brs": "thewalrus_b200.brs(A, E)"}[kind] + " with a host NumPy array"
        line = {
            "metric": metric_name(args.workload),
            "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "n": n, "units_per_step": units,
                       "units": ("photon-number patterns" if is_gbs else "photon-number samples (accepted or not)"
                                 if kind == "hsample" else "subsets (reference `steps`)"),
                       "input": cfg_in,
                       "parallelism": (f"pattern shards x{world}, one all-gather" if is_gbs else f"subset-index shards x{world}, one all-reduce"),
                       "l2_flush": True,
                       "l2_note": "256 MiB fill between timed iterations; inputs are KB-sized, the path is FP64-pipe bound"},
            "e2e": {"value": units * args.steps / e2e_s, "unit": unit, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / args.steps * 1e3, "api": api},
            "gpu_launches": launches_per_step * args.steps * world,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "pipe": pipe,
                         "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s", "frac": achieved / peak.value,
                         "peak_source": "measured live by wb200_fp64_peak (dependent-chain micro-benchmark on the same pipe: "
                                        "DMMA for the hafnian kernels, DFMA otherwise); MEASURED_PEAKS.json has no FP64 entry; "
                                        "nominal 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2",
                         "flops_per_unit": my_flops, "flops_model": model,
                         "reference_algorithm_equivalent_tflops": per_gpu_units * ref_flops / (kms * 1e-3) * 1e-12,
                         "kernel_ms": kms, "traffic": MEASURED_TRAFFIC.get(args.workload),
                         "traffic_note": "bytes per launch of the dominant kernel, dram__bytes_read.sum + dram__bytes_write.sum from "
                                         "the ncu captures under profiles/ (r01_ncu_traffic_hafnian50.csv, r01_ncu_*.txt); the "
                                         "algorithmic traffic is KB (matrix in, 4 doubles per CTA out) — these paths are FP64-pipe "
                                         "bound, DRAM sees the 4-byte spill slot and L2 write-backs over a multi-second launch"},
            "result": {"re": res.real, "im": res.imag, "e2e_re": complex(r_e2e).real},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(kind, n, X, cutoff=args.cutoff)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
