"""GPU twins of the caller tests (tests/test_quantum_callers.py, tests/test_moments.py): the same comparisons against
outputs of the reference, through the CUDA kernels instead of the oracle.  Kept in a file that sorts last: these
were written after the last GPU run of round 1, so under ``-x`` they cannot mask the suite that was validated."""
import pytest

import test_moments as tm
import test_quantum_callers as tq
from test_moments import gold as gold_moments  # noqa: F401  (fixtures)
from test_quantum_callers import gold, gold_ft, gold_marg  # noqa: F401

pytestmark = pytest.mark.gpu


def test_gpu_pure_state_callers_vs_reference(gold):  # noqa: F811
    tq._check_pure(gold)


def test_gpu_density_matrix_vs_reference(gold):  # noqa: F811
    tq._check_mixed(gold)


def test_gpu_fock_tensor_vs_reference(gold_ft):  # noqa: F811
    tq._check_fock_tensor(gold_ft)


def test_gpu_marginals_and_tvd_bounds_vs_reference(gold_marg):  # noqa: F811
    tq._check_marginals(gold_marg)


def test_gpu_moments_vs_reference(gold_moments):  # noqa: F811
    tm._check(gold_moments)
