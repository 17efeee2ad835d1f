import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, ".."))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(REPO, "i-adhore_b200", "python"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libs():
    """The oracle (CPU checker) and the product library must exist; build what is missing.
    nvcc cross-compiles sm_100a without a GPU."""
    from iadhore_b200 import capi
    if not os.path.exists(capi.ORACLE_LIB):
        subprocess.run(["make", "-C", os.path.join(REPO, "oracle"), "liboracle.so"], check=True)
    if not os.path.exists(capi.PRODUCT_LIB):
        subprocess.run(["make", "-C", os.path.join(REPO, "i-adhore_b200", "csrc")], check=True)
    yield
