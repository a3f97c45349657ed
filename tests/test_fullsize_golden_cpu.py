"""CPU checks of the full-size golden file (tests/golden/reference_fullsize.json) and of the host-side exact path.

* the reference's own outputs and the oracle's agree with each other inside the golden (the oracle is pinned at full
  size, not only on the small cases of test_oracle_golden.py);
* the exact perfect-matching counts (thewalrus.reference.hafnian / int64 recursive_hafnian of the reference) are
  reproduced by the package's exact host recursion, which is what `hafnian(A, method="recursive")` returns for integer
  matrices."""
import json
import os
import sys

import numpy as np
import pytest
from conftest import ROOT, rel

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden_fullsize as mg  # noqa: E402  (input recipes only)

import thewalrus_b200 as wb  # noqa: E402
from thewalrus_b200._hafnian import recursive_hafnian  # noqa: E402


@pytest.fixture(scope="module")
def full():
    with open(os.path.join(ROOT, "tests", "golden", "reference_fullsize.json")) as fh:
        return json.load(fh)


def cz(d):
    return complex(d["re"], d["im"])


def test_reference_and_oracle_agree_inside_the_golden(full):
    e = full["hafnian24"]
    assert rel(cz(e["reference"]), cz(e["oracle_ld"])) < 1e-12 and rel(cz(e["reference_loop"]), cz(e["oracle_ld_loop"])) < 1e-12
    e = full["tor48"]
    assert rel(cz(e["reference_rec"]).real, e["oracle_ld"]) < 1e-11 and rel(e["oracle_double"], e["oracle_ld"]) < 1e-11
    assert rel(cz(e["ltor_oracle_double"]), cz(e["ltor_oracle_ld"])) < 1e-11
    if "perm32" in full:
        e = full["perm32"]
        # the reference's plain-FP64 bbfg sum of 2^31 terms is itself only ~5e-8 from the long-double value (kappa ~ 1e7)
        assert rel(cz(e["oracle_double"]), cz(e["oracle_ld"])) < 1e-7      # 6e-9 measured: plain FP64 sum in the port too
        if "reference_bbfg" in e:
            assert rel(cz(e["reference_bbfg"]), cz(e["oracle_ld"])) < 1e-6
    if "gbs16" in full and "sample_reference" in full["gbs16"]:
        e = full["gbs16"]
        ld = np.load(os.path.join(ROOT, "tests", "golden", "gbs16_probabilities_ld.npy"))
        ref = np.array(e["sample_reference"])
        got = ld[np.array(e["sample_index"])]
        assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3 * np.max(np.abs(ref)))) < 1e-10
        assert rel(e["oracle_double_sum"], e["oracle_ld_sum"]) < 1e-12
    if "hafnian50" in full and "oracle_double" in full["hafnian50"]:
        e = full["hafnian50"]
        for w in e["windows_ld"]:
            assert rel(cz(w["double"]), cz(w["ld"])) < 1e-11
        if "reference" in e:
            assert rel(cz(e["reference"]), cz(e["oracle_double"])) < 1e-9


def test_exact_counts_by_the_host_recursion(full):
    for e in full["exact"]:
        if e["n"] > 24:
            continue
        A = mg.er_graph(e["n"], e["p"], e["seed"])
        assert np.allclose(mg.fp(A), e["fp"])
        v = recursive_hafnian(A)
        assert isinstance(v, np.integer) and int(v) == e["hafnian"], (e["n"], v, e["hafnian"])
        got = wb.hafnian(A, method="recursive")           # int64 in, exact int out: no GPU involved
        assert isinstance(got, np.integer) and int(got) == e["hafnian"]


def test_recursive_hafnian_matches_the_oracle_on_complex_input():
    from oracle import walrus_oracle as wo

    rng = np.random.default_rng(3)
    for n in (2, 4, 6, 10):
        G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        A = G + G.T
        assert rel(recursive_hafnian(A), wo.haf(A)) < 1e-12
    assert recursive_hafnian(np.ones((5, 5), dtype=np.int64)) == 0 and recursive_hafnian(np.zeros((0, 0))) == 1
