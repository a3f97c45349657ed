import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _have_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a box without CUDA; `-m gpu` on the B200 box runs them all."""
    if _have_cuda():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (run on the B200 box: pytest -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_outputs.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden_next():
    """Reference outputs for the SURVEY 8(f) components (tests/golden/make_golden_next.py)."""
    with open(os.path.join(ROOT, "tests", "golden", "reference_outputs_next.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the CUDA library and the oracle are built (nvcc cross-compiles without a GPU)."""
    import shutil

    import __graft_entry__ as ge

    if shutil.which(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")) or os.path.exists(ge.LIB):
        ge.build()
    else:   # no CUDA toolchain and no prebuilt library: the oracle / host-logic tests still run
        from oracle import build as obuild

        obuild.ensure()


def dec(d):
    return np.array(d["re"]) + 1j * np.array(d["im"])


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def random_symmetric(rng, n, kind="complex"):
    if kind == "complex":
        G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    else:
        G = rng.standard_normal((n, n))
    return G + G.T
