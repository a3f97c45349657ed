"""north_star target under torchrun (NCCL, N ranks): hafnian of 56 x 56 and permanent of 40 x 40 matrices sharded over
the ranks, COMPLETE results against the long-double goldens (tests/golden/reference_fullsize.json), wall time per call.
torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/gpu_fullsize_x8.py"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import bench  # noqa: E402
import make_golden_fullsize as mg  # noqa: E402
import thewalrus_b200 as wb  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
full = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_fullsize.json")))


def cz(d):
    return complex(d["re"], d["im"])


def run(name, fn, want):
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    got = fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([complex(got).real, complex(got).imag], dtype=torch.float64, device="cuda")
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        err = abs(got - want) / abs(want)
        print(f"[x{world}] {name}: {dt * 1e3:.1f} ms  rel_err vs golden {err:.3e}  spread across ranks {float((hi - lo).abs().max()):.1e}  value {got!r}",
              flush=True)
        return err
    return 0.0


errs = []
for e in full["structured"]:
    if e["kind"] == "haf_bipartite" and e["n"] == 56:
        _, A = mg.bipartite_input(e["n"], e["seed"])
        errs.append(run("hafnian 56x56, bipartite (= long-double perm 28x28)", lambda: wb.hafnian(A, group=True), cz(e["value"])))
    if e["kind"] == "haf_direct_sum" and e["n"] == 56:
        _, A2 = mg.direct_sum_input(e["n1"], e["n2"], e["seed"])
        errs.append(run("hafnian 56x56, direct sum 28 + 28", lambda: wb.hafnian(A2, group=True), cz(e["value"])))
        errs.append(run("loop hafnian 56x56, direct sum 28 + 28", lambda: wb.hafnian(A2, loop=True, group=True), cz(e["loop_value"])))
    if e["kind"] == "perm_blocks" and e["n"] == 40:
        _, M = mg.block_perm_input(e["n1"], e["n2"], e["seed"])
        errs.append(run(f"permanent 40x40, blocks {e['n1']} + {e['n2']}", lambda: wb.perm(M, method="glynn", group=True), cz(e["value"])))
_, _, U = bench.make_input("perm32")
errs.append(run("permanent 32x32 (BASELINE config 2 input)", lambda: wb.perm(U, method="glynn", group=True), cz(full["perm32"]["oracle_ld"])))
_, _, O = bench.make_input("tor48")
errs.append(run("torontonian 2N = 48 (BASELINE config 4 input)", lambda: wb.tor(O, group=True), full["tor48"]["oracle_ld"]))
if "hafnian50" in full and "oracle_double" in full["hafnian50"]:
    _, _, A50 = bench.make_input("hafnian50")
    errs.append(run("hafnian 50x50 (metric input)", lambda: wb.hafnian(A50, group=True), cz(full["hafnian50"]["oracle_double"])))
if rank == 0:
    print("FULLSIZE x%d" % world, "OK" if max(errs) < 1e-10 else "FAIL", "worst rel_err %.3e" % max(errs))
dist.destroy_process_group()
