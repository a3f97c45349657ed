"""CPU twin of the reference test-suite run (tests/golden/ref_suite): the reference's own test files against the
package's host logic with the CPU oracle answering the kernel calls (tests/oracle_engine.py).  The `-m gpu` run on
the B200 box executes the same files through the CUDA kernels."""
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_test_files_pass_on_host_logic_with_oracle_kernels():
    env = dict(os.environ, WB200_REFSUITE_CPU="1")
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "golden", "ref_suite"), "-q", "-x",
                          "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    tail = res.stdout[-2000:] + res.stderr[-2000:]
    assert res.returncode == 0, tail
    assert " passed" in res.stdout and "failed" not in res.stdout, tail
