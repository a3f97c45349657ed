"""Small-n hafnian sweep: kernel-only ms per call (CUDA events around wb200_hafnian_dev, 200 calls) and end-to-end ms
through thewalrus_b200.hafnian, for the panel-split launch shapes against the round-1 shapes (WB200_HAF_WARPS)."""
import os
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch

import bench
import thewalrus_b200 as wb
from thewalrus_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream(dev)
print("n   mode      kernel_ms   e2e_ms   value")
for n in (8, 12, 16, 20, 22, 24, 26, 28, 30, 32, 34):
    rng = np.random.default_rng(n)
    G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    A = G + G.T
    x, _, _ = wb.matched_reps([1] * n)
    Ax = np.ascontiguousarray(A[np.ix_(x, x)])
    dA = torch.from_numpy(Ax.view(np.float64).reshape(-1)).to(dev)
    ws = torch.empty(lib.wb200_hafnian_workspace_bytes(n) // 8 + 8, dtype=torch.float64, device=dev)
    out = torch.zeros(4, dtype=torch.float64, device=dev)
    for mode in ("new", "r01"):
        if mode == "r01":
            steps = 1 << (n // 2 - 1)
            g = (steps + 3) // 4
            os.environ["WB200_HAF_WARPS"] = "4" if g <= 4 * 148 else ("8" if g <= 8 * 148 else "12")
        else:
            os.environ.pop("WB200_HAF_WARPS", None)
        reps = 200
        for _ in range(5):
            lib.wb200_hafnian_dev(dA.data_ptr(), None, n, 0, 1 << (n // 2 - 1), out.data_ptr(), ws.data_ptr(), ws.numel() * 8, st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            lib.wb200_hafnian_dev(dA.data_ptr(), None, n, 0, 1 << (n // 2 - 1), out.data_ptr(), ws.data_ptr(), ws.numel() * 8, st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        kms = e0.elapsed_time(e1) / reps
        v = wb.hafnian(A)
        t0 = time.perf_counter()
        for _ in range(reps):
            v = wb.hafnian(A)
        e2e = (time.perf_counter() - t0) / reps * 1e3
        print(f"{n:3d} {mode:5s} {kms:10.4f} {e2e:9.4f}   {v:.12g}")
