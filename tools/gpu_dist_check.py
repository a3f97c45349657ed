"""Multi-GPU parity under torchrun (NCCL): every sharded entry point against its single-GPU value.
torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/gpu_dist_check.py"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import thewalrus_b200 as wb
import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(5)
def rel(a, b): return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(np.asarray(b)), 1e-300)))
checks = {}
n = 30
G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)); A = G + G.T
checks["hafnian30"] = rel(wb.hafnian(A, group=True), wb.hafnian(A))
checks["lhaf30"] = rel(wb.hafnian(A, loop=True, group=True), wb.hafnian(A, loop=True))
A8 = A[:8, :8] / 3
rpt = [2, 1, 0, 3, 1, 1, 2, 2]
checks["hafnian_repeated"] = rel(wb.hafnian_repeated(A8, rpt, mu=np.diag(A8), loop=True, group=True),
                                 wb.hafnian_repeated(A8, rpt, mu=np.diag(A8), loop=True))
_, _, U = bench.make_input("perm25")
checks["perm25_glynn"] = rel(wb.perm(U, method="glynn", group=True), wb.perm(U, method="glynn"))
checks["perm25_ryser"] = rel(wb.perm(U, method="ryser", group=True), wb.perm(U, method="ryser"))
Mi = rng.integers(-3, 4, (14, 14)).astype(np.int64)
checks["perm_int64"] = abs(wb.perm(Mi, "ryser", group=True) - wb.perm(Mi, "ryser"))
_, _, O = bench.make_input("tor28")
checks["tor28"] = rel(wb.tor(O, group=True), wb.tor(O))
A5 = A[:5, :5] / 4; D5 = np.diag(A)[:5] / 4
checks["lhaf_batch"] = rel(wb.loop_hafnian_batch(A5, D5, [1, 0, 2, 1], 6, group=True), wb.loop_hafnian_batch(A5, D5, [1, 0, 2, 1], 6))
mu, cov, pats = bench.make_gbs_state(8, 1001, seed=77)
checks["probabilities_batch"] = rel(wb.probabilities_batch(mu, cov, pats, group=True) + 1e-300, wb.probabilities_batch(mu, cov, pats) + 1e-300)
_, _, (Ol, gl) = bench.make_input("ltor28")
checks["ltor28"] = rel(wb.ltor(Ol, gl, group=True), wb.ltor(Ol, gl))
_, _, Am = bench.make_input("mtl10")
checks["mtl10"] = rel(wb.mtl(Am, group=True), wb.mtl(Am))
zeta = 0.3 * (rng.standard_normal(20) + 1j * rng.standard_normal(20))
checks["lmtl10"] = rel(wb.lmtl(Am, zeta, group=True), wb.lmtl(Am, zeta))
_, _, (Ab, Eb) = bench.make_input("brs9")
checks["brs9"] = rel(wb.brs(Ab, Eb, group=True), wb.brs(Ab, Eb))
Dg = rng.standard_normal((3, 5)) + 1j * rng.standard_normal((3, 5))
checks["lhaf_batch_gamma"] = rel(wb.loop_hafnian_batch_gamma(A5, Dg, [1, 0, 2, 1], 5, group=True),
                                 wb.loop_hafnian_batch_gamma(A5, Dg, [1, 0, 2, 1], 5))
# all ranks must hold bit-identical results
v = wb.hafnian(A, group=True)
t = torch.tensor([v.real, v.imag], dtype=torch.float64, device="cuda")
lo, hi = t.clone(), t.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
checks["bit_identical_across_ranks"] = float((hi - lo).abs().max().item())
if rank == 0:
    print("dist checks (world %d):" % world, {k: float(v) for k, v in checks.items()})
    bad = {k: v for k, v in checks.items() if not v < 1e-10}
    print("DIST PARITY", "FAIL %s" % bad if bad else "OK")
dist.destroy_process_group()
