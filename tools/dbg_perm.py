import json, os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import thewalrus_b200 as wb
from oracle import c_oracle as co
G = json.load(open("tests/golden/reference_outputs.json"))
def dec(d): return np.array(d["re"]) + 1j * np.array(d["im"])
def rel(a, b): return abs(a - b) / max(abs(b), 1e-300)
for c in G["perm"]:
    if c["kind"] == "int": continue
    A = dec(c["A"]); A = A.real if c["kind"] == "real" else A
    exact = co.perm(A, "bbfg", long_double=True)
    for method in ("bbfg", "ryser"):
        got = wb.perm(A, method)
        print(c["n"], c["kind"], method, "gpu-vs-ref %.2e gpu-vs-exact %.2e ref-vs-exact %.2e" % (rel(got, dec(c[method])), rel(got, exact), rel(dec(c[method]), exact)))
