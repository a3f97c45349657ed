#!/usr/bin/env python
"""Headline benchmark: BASELINE.json metric "hafnian n=50 complex128 subsets/s & wall-time at 1/2/4/8 B200;
% FP64 peak".

One "step" = one complete hafnian of a random 50x50 complex128 symmetric matrix (2^24 Glynn subsets,
seed 1000*1+50 as in SURVEY.md 8d).  With N GPUs the subset index is sharded in contiguous ranges (strong
scaling: total work fixed) and the partial sums are combined by one all-reduce inside the timed region.

  python bench.py --gpus N --steps K --warmup W                 # GPU arm (torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K --warmup W  # CPU arm: the oracle port on host cores

Other workloads for development: --workload hafnian56|hafnian40|perm32|perm36|tor48|lhaf50
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
    if kind == "tor":  # n = 2N
        N = n // 2
        rng = np.random.default_rng(1000 * 4 + n)
        B = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        H = B @ B.conj().T
        return kind, n, 0.9 * H / np.linalg.norm(H, 2)
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


def units_and_flops(kind, n):
    """(units per step, algorithmic flops per unit of THIS implementation, reference-algorithm flops per unit)."""
    if kind in ("hafnian", "lhaf"):
        m = n // 2
        nprod = (m - 1) // 2
        return 1 << (m - 1), 8.0 * n**3 * nprod, 8.0 * n**3 * (m - 1)
    if kind == "perm":
        return 1 << (n - 1), 8.0 * n - 4, 8.0 * n - 4
    if kind == "tor":
        N = n // 2
        return 1 << N, float("nan"), float("nan")
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


def cpu_baseline(kind, n, X, seconds=15.0):
    """Time the oracle's C port (reference algorithm) on the host cores over a bounded sample."""
    from oracle import c_oracle as co

    threads = co.max_threads()
    if kind in ("hafnian", "lhaf"):
        x = co.matched_order(X)
        Ax = np.ascontiguousarray(X[np.ix_(x, x)])
        Dx = np.ascontiguousarray(np.diag(X)[x]) if kind == "lhaf" else None
        co.hafnian_range(Ax, 0, 64 * threads, Dx)  # warm-up (thread pool, page faults)
        t0 = time.perf_counter()
        co.hafnian_range(Ax, 0, 256 * threads, Dx)
        rate = 256 * threads / (time.perf_counter() - t0)
        sample = int(min(1 << (n // 2 - 1), max(1024, rate * seconds)))
        t0 = time.perf_counter()
        co.hafnian_range(Ax, 0, sample, Dx)
        dt = time.perf_counter() - t0
        what = f"first {sample} of {1 << (n // 2 - 1)} Glynn subsets of the same {n}x{n} matrix, full product-chain algorithm"
    elif kind == "perm":
        sample = int(min(1 << (n - 1), 4e7 * threads))
        t0 = time.perf_counter()
        co.perm_range(X, 0, 0, sample)
        dt = time.perf_counter() - t0
        what = f"first {sample} of {1 << (n - 1)} Gray-code steps (the reference itself is single-threaded; the port splits the range over threads)"
    else:
        N = n // 2
        sample = 1 << N
        t0 = time.perf_counter()
        co.tor_recursive(X)
        dt = time.perf_counter() - t0
        threads = 1
        what = "full recursive torontonian (single thread, as the reference)"
    return {"value": sample / dt, "unit": "subsets/s", "cores": threads, "kind": "port", "sample": what,
            "seconds": dt}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as ge  # builds the oracle if needed (CPU only)

    from oracle import build as obuild

    obuild.ensure()
    kind, n, X = make_input(args.workload)
    units, _, ref_flops = units_and_flops(kind, n)
    vals = []
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(kind, n, X, seconds=max(2.0, 60.0 / max(1, args.warmup + args.steps)))
        if i >= args.warmup:
            vals.append(cb)
    v = statistics.mean(c["value"] for c in vals)
    line = {"impl": "reference", "metric": METRIC if args.workload == "hafnian50" else f"{args.workload} subsets/s",
            "value": v, "unit": "subsets/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": units / v * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "n": n, "units_per_step": units,
                       "note": "each step is a bounded sample of the workload; ms_per_step is extrapolated to the full step"},
            "cpu_baseline": {"value": v, "unit": "subsets/s", "cores": vals[-1]["cores"], "kind": "port",
                             "sample": vals[-1]["sample"], "gflops_reference_algorithm": v * ref_flops * 1e-9},
            "e2e": {"value": v, "unit": "subsets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="hafnian50")
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
    kind, n, X = make_input(args.workload)
    units, my_flops, ref_flops = units_and_flops(kind, n)
    lo, hi = shard_range(units if kind != "tor" else _engine.tor_num_prefixes(n // 2), rank, world)

    # ---- device-resident inputs for the kernel-only number
    if kind in ("hafnian", "lhaf"):
        x, er, _ = wb.matched_reps([1] * n)
        Ax = np.ascontiguousarray(X[np.ix_(x, x)].astype(np.complex128))
        dA = torch.from_numpy(Ax.view(np.float64).reshape(-1)).to(dev)
        dD = torch.from_numpy(np.ascontiguousarray(np.diag(X)[x]).view(np.float64).reshape(-1)).to(dev) if kind == "lhaf" else None
        wsb = lib.wb200_hafnian_workspace_bytes(n)
        launches_per_step = 3
    elif kind == "perm":
        dA = torch.from_numpy(np.ascontiguousarray(X).view(np.float64).reshape(-1)).to(dev)
        wsb = lib.wb200_perm_workspace_bytes(n)
        launches_per_step = 2
    else:
        dA, wsb, launches_per_step = None, 8, 3
    ws = torch.empty((wsb + 7) // 8, dtype=torch.float64, device=dev)
    out = torch.zeros(4, dtype=torch.float64, device=dev)
    table = torch.zeros((world, 4), dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream(dev)

    def kernel_step():
        if kind in ("hafnian", "lhaf"):
            rc = lib.wb200_hafnian_dev(dA.data_ptr(), dD.data_ptr() if dD is not None else None, n, lo, hi, out.data_ptr(),
                                       ws.data_ptr(), ws.numel() * 8, stream.cuda_stream)
        elif kind == "perm":
            rc = lib.wb200_perm_dev(dA.data_ptr(), n, 0, lo, hi, out.data_ptr(), ws.data_ptr(), ws.numel() * 8,
                                    stream.cuda_stream)
        else:
            o2 = _engine.tor_range(X, lo, hi, dev)
            out[:2] = torch.from_numpy(o2).to(dev)
            rc = 0
        _lib.check(rc, "kernel step")
        if world > 1:  # the one collective of the path: all-reduce of the (hi, lo) partials
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
        flush.fill_(1.0)  # L2 flush between timed iterations (inputs are 40 KB, far below L2)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        kernel_step()
        e1.record(stream)
        sync_all()
        ms = e0.elapsed_time(e1)
        kern_ms.append(ms)
        total_ms += ms
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = units * args.steps / (total_ms * 1e-3)
    if world > 1:
        res = _engine.combine4(table.cpu().numpy())
    else:
        res = _engine.combine4([out.cpu().numpy()])

    # ---- end to end through the public API with host buffers
    def e2e_step():
        if kind == "hafnian":
            return wb.hafnian(X, group=(True if world > 1 else None))
        if kind == "lhaf":
            return wb.hafnian(X, loop=True, group=(True if world > 1 else None))
        if kind == "perm":
            return wb.perm(X, method="glynn", group=(True if world > 1 else None))
        return wb.tor(X, group=(True if world > 1 else None))

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
        lib.wb200_fp64_peak(local, 1, ctypes.byref(peak))
        per_gpu_units = (hi - lo)
        kms = statistics.mean(kern_ms)
        achieved = per_gpu_units * my_flops / (kms * 1e-3) * 1e-12
        line = {
            "metric": METRIC if args.workload == "hafnian50" else f"{args.workload} subsets/s",
            "value": value, "unit": "subsets/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "n": n, "units_per_step": units, "units": "Glynn subsets (reference `steps`)",
                       "input": "random complex symmetric G+G^T, seed 1000*config+n" if kind != "perm" else "n x n block of a 2n Haar unitary",
                       "parallelism": f"subset-index shards x{world}, one all-reduce", "l2_flush": True,
                       "l2_note": "256 MiB fill between timed iterations; inputs are KB-sized, the path is FP64-pipe bound"},
            "e2e": {"value": units * args.steps / e2e_s, "unit": "subsets/s", "h2d_bytes_per_step": int(X.nbytes),
                    "d2h_bytes_per_step": 32, "ms_per_step": e2e_s / args.steps * 1e3,
                    "api": "thewalrus_b200.hafnian(A) with a host NumPy array" if kind == "hafnian" else "public API"},
            "gpu_launches": launches_per_step * args.steps * world,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "pipe": "FP64 DMMA.8x8x4 (same issue rate as the FP64 FMA pipe)",
                         "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s", "frac": achieved / peak.value,
                         "peak_source": "measured live by wb200_fp64_peak (DMMA chain micro-benchmark); MEASURED_PEAKS.json has no FP64 entry; nominal 37.2 at 1965 MHz",
                         "flops_per_unit": my_flops, "flops_model": "8 n^3 floor((n/2-1)/2): pairing halves the reference's 8 n^3 (n/2-1)",
                         "reference_algorithm_equivalent_tflops": per_gpu_units * ref_flops / (kms * 1e-3) * 1e-12,
                         "kernel_ms": kms, "traffic": None},
            "result": {"re": res.real, "im": res.imag, "e2e_re": complex(r_e2e).real},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(kind, n, X)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
