"""Ad-hoc GPU sanity run: golden parity + first timings (development aid, not the test-suite)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import thewalrus_b200 as wb
from thewalrus_b200 import _engine, _lib
from oracle import walrus_oracle as wo

G = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "reference_outputs.json")))
def dec(d): return np.array(d["re"]) + 1j * np.array(d["im"])
def rel(a, b): return abs(a - b) / max(abs(b), 1e-300)
worst = {}
def upd(k, e, tag=None):
    if e > worst.get(k, (0, None))[0]: worst[k] = (float(e), tag)
    worst.setdefault(k, (0.0, None))
t0 = time.time()
for c in G["hafnian"]:
    A = dec(c["A"]); A = A.real if c["kind"] == "real" else A
    upd("haf_glynn", rel(wb.hafnian(A), dec(c["glynn"])), c["n"])
    if c["n"] <= 16: upd("haf_inclexcl", rel(wb.hafnian(A, method="inclexcl"), dec(c["inclexcl"])), c["n"])
for c in G["loop_hafnian"]:
    A = dec(c["A"]); A = A.real if c["kind"] == "real" else A
    upd("lhaf", rel(wb.hafnian(A, loop=True), dec(c["value"])), c["n"])
for c in G["hafnian_repeated"]:
    upd("haf_rep", rel(wb.hafnian_repeated(dec(c["A"]), c["rpt"], glynn=c["glynn"]), dec(c["value"])), str(c["rpt"]))
for c in G["loop_hafnian_reps"]:
    upd("lhaf_rep", rel(wb.loop_hafnian(dec(c["A"]), dec(c["mu"]), c["rpt"], glynn=c["glynn"]), dec(c["value"])), str(c["rpt"]))
for c in G["perm"]:
    if c["kind"] == "int":
        A = np.array(c["A"], dtype=np.int64)
        upd("perm_int_ryser", abs(wb.perm(A, "ryser") - c["ryser"]), c["n"]); upd("perm_int_bbfg", abs(wb.perm(A, "bbfg") - c["bbfg"]), c["n"])
    else:
        A = dec(c["A"]); A = A.real if c["kind"] == "real" else A
        upd("perm_bbfg_" + c["kind"], rel(wb.perm(A, "bbfg"), dec(c["bbfg"])), c["n"]); upd("perm_ryser_" + c["kind"], rel(wb.perm(A, "ryser"), dec(c["ryser"])), c["n"])
for c in G["tor"]:
    O = dec(c["O"]); O = O.real if c["kind"] == "real" else O
    upd("tor_" + c["kind"], rel(wb.tor(O), dec(c["rec"])), c["N"])
for c in G["int_hafnian"]:
    A = np.array(c["A"], dtype=np.float64)
    upd("int_haf_abs", abs(wb.hafnian(A).real - c["value"]), c["n"])
print("parity (worst rel err, case):", json.dumps(worst, indent=1), "in %.1fs" % (time.time() - t0))

# timings through the host C-ABI
import ctypes
lib = _lib.load()
rng = np.random.default_rng(1)
for n in (24, 32, 40):
    Gm = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)); A = Gm + Gm.T
    x, er, _ = wb.matched_reps([1] * n); Ax = np.ascontiguousarray(A[np.ix_(x, x)])
    out = np.zeros(4); ms = ctypes.c_double(0)
    steps = 1 << (n // 2 - 1)
    for rep in range(2):
        rc = lib.wb200_hafnian_host(0, _lib.dptr(Ax.view(np.float64)), None, n, 0, steps, _lib.dptr(out), ctypes.byref(ms))
    m = n // 2; T = (m + 3) // 4; nprod = (m - 1) // 2
    dmma = steps / 4 * m * nprod * (2 * T * T * 4)
    print("hafnian n=%d steps=%d kernel %.3f ms -> %.3e subsets/s, executed %.2f TFLOP/s (DMMA), ref-alg %.2f TFLOP/s" % (
        n, steps, ms.value, steps / ms.value * 1e3, dmma * 512 / ms.value * 1e-9, steps * 8.0 * n**3 * (m - 1) / ms.value * 1e-9))
for n in (24, 28, 32):
    U = np.linalg.qr(rng.standard_normal((2 * n, 2 * n)) + 1j * rng.standard_normal((2 * n, 2 * n)))[0][:n, :n]
    U = np.ascontiguousarray(U); out = np.zeros(4); ms = ctypes.c_double(0)
    steps = 1 << (n - 1)
    for rep in range(2):
        rc = lib.wb200_perm_host(0, _lib.dptr(U.view(np.float64)), n, 0, 0, steps, _lib.dptr(out), ctypes.byref(ms))
    print("perm n=%d kernel %.3f ms -> %.3e subsets/s, %.2f TFLOP/s (8n-4 flops/subset)" % (n, ms.value, steps / ms.value * 1e3, steps * (8.0 * n - 4) / ms.value * 1e-9))
for N in (12, 16, 20, 24):
    B = rng.standard_normal((2 * N, 2 * N)) + 1j * rng.standard_normal((2 * N, 2 * N)); H = B @ B.conj().T
    O = np.ascontiguousarray(0.9 * H / np.linalg.norm(H, 2)); out = np.zeros(2); ms = ctypes.c_double(0)
    tot = _engine.tor_num_prefixes(N)
    for rep in range(2):
        rc = lib.wb200_tor_host(0, _lib.dptr(O.view(np.float64)), N, 0, tot, _lib.dptr(out), ctypes.byref(ms))
    print("tor N=%d kernel %.3f ms -> %.3e subsets/s value %.6e" % (N, ms.value, 2.0**N / ms.value * 1e3, out[0] + out[1]))
