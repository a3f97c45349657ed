"""One hafnian call of size n (default 50) through the public API: the target of the ncu DRAM-traffic capture."""
import sys

sys.path.insert(0, ".")
import bench
import thewalrus_b200 as wb

w = sys.argv[1] if len(sys.argv) > 1 else "hafnian50"
kind, n, A = bench.make_input(w)
print(w, wb.hafnian(A) if kind == "hafnian" else (wb.perm(A) if kind == "perm" else wb.tor(A)))
