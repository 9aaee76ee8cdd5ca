import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle behind the same C-ABI structs as the product (tests only)."""
    from rtg_tools_b200 import capi
    so = os.path.join(ROOT, "oracle", "liboracle.so")
    src = os.path.join(ROOT, "oracle", "vcfeval_oracle.cpp")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return capi.Library(so, prefix="oracle_")


@pytest.fixture(scope="session")
def fixtures():
    with open(os.path.join(ROOT, "tests", "golden", "vcfeval_fixtures.json")) as f:
        return json.load(f)["scenarios"]
