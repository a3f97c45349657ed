"""bench.py contract checks that run without a GPU: the reference arm prints one JSON line with the agreed keys,
non-zero ranks of a torchrun launch stay silent, and the workload table is self-consistent."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e,
                         timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return [ln for ln in out.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_line():
    lines = _run(["--impl", "reference", "--workload", "hafnian24", "--steps", "1", "--warmup", "0"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["metric"] == "hafnian24 subsets/s" and d["unit"] == "subsets/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    if d["cpu_baseline"]["kind"] == "port":      # numba reference not importable here: the line must say why
        assert "reference_unavailable" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "hafnian24" and d["config"]["units_per_step"] == 2048
    # the reference arm prints the SAME config keys as the GPU arm (the driver's same_config check)
    args = type("A", (), {"gpus": 1, "batch": 100000, "cutoff": 6})()
    want = bench.describe("hafnian24", args)[4]
    assert {k: d["config"][k] for k in want} == want


def test_reference_arm_port_fallback_line():
    lines = _run(["--impl", "reference", "--workload", "perm12", "--steps", "1", "--warmup", "0", "--seconds", "0.2"])
    d = json.loads(lines[0])
    assert d["cpu_baseline"]["kind"] == "port" and "reference_unavailable" in d["cpu_baseline"] and d["value"] > 0


def test_reference_arm_other_ranks_are_silent():
    assert _run(["--impl", "reference", "--workload", "hafnian24", "--steps", "1", "--warmup", "0"],
                env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_headline_metric_and_inputs():
    assert bench.metric_name("hafnian50") == "hafnian n=50 complex128 subsets/s"
    kind, n, A = bench.make_input("hafnian50")
    assert (kind, n) == ("hafnian", 50) and A.shape == (50, 50) and np.allclose(A, A.T) and np.iscomplexobj(A)
    units, mine, ref, _, _ = bench.units_and_flops("hafnian", 50)
    # symmetric-half kernel: 1444 of the 2500 entries of each of the 12 products (hafnian_sym.cu)
    assert units == 1 << 24 and ref == 8 * 50**3 * 24 and mine == 8 * 50 * 1444 * 12
    assert bench.haf_sym_entries(48) == 1344 and bench.haf_sym_entries(56) == 1792 and bench.haf_sym_entries(34) == 0 and bench.haf_sym_entries(64) == 2304 and bench.haf_sym_entries(40) == 960
    assert bench.haf_sym_entries(46) == 1236 < bench.haf_sym_entries(48)       # padded to the 48 shape: only real entries count
    assert bench.units_and_flops("lhaf", 50)[1] == 8 * 50**3 * 12            # loops stay on the row-panel kernel
    assert bench.units_and_flops("hafnian", 56)[1] == 8 * 56 * 1792 * 13 and bench.units_and_flops("hafnian", 34)[1] == 8 * 34**3 * 8
    kind, n, U = bench.make_input("perm32")
    assert U.shape == (32, 32) and np.linalg.norm(U, 2) <= 1 + 1e-12          # block of a unitary
    assert bench.units_and_flops("perm", 32)[0] == 1 << 31
    kind, n, O = bench.make_input("tor48")
    assert O.shape == (48, 48) and np.allclose(O, O.conj().T) and np.all(np.linalg.eigvalsh(np.identity(48) - O) > 0)
    mu, cov, pats = bench.make_gbs_state(16, 1000, seed=3016)
    assert pats.shape == (1000, 16) and pats.sum(axis=1).max() <= 10 and cov.shape == (32, 32)


def test_gbs_flop_models_are_consistent():
    """Executed (useful) flops of the batched kernels are below the reference-algorithm equivalent (trace pairing and
    un-expanded mixed-radix subsets) and scale with the number of patterns."""
    M, mu, cov, pats, A, gamma, rpt = bench.gbs_inputs("gbs16", 2000)
    ex, ref = bench.gbs_executed_flops(rpt), bench.gbs_reference_flops(pats)
    assert 0 < ex < ref
    assert np.isclose(bench.gbs_executed_flops(np.concatenate([rpt, rpt])), 2 * ex)
