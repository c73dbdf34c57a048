import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # everything is built in-tree by the top-level Makefile (seconds; nvcc cross-compiles sm_100a)
    r = subprocess.run(["make", "-C", ROOT, "all"], capture_output=True, text=True)
    if r.returncode != 0:
        raise pytest.UsageError("make all failed:\n" + r.stdout[-3000:] + r.stderr[-3000:])


@pytest.fixture(scope="session")
def oracle():
    """The checker of the parity tests: the unmodified reference (oracle/_ref) where it is present, else the C restatement."""
    from tests import oracle_lib
    return oracle_lib.BestOracle()


@pytest.fixture(scope="session")
def port():
    """The C restatement itself (oracle/dsd_oracle.c), for the tests that pin it to the reference and the golden vectors."""
    from tests import oracle_lib
    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def ref():
    from tests import oracle_lib
    if not os.path.exists(oracle_lib.REF_SO):
        pytest.skip("oracle/_ref/libdsdref.so not built (no /root/reference on this box)")
    return oracle_lib.Oracle(reference=True)
