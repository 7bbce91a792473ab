import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
BIOD_DATA = os.path.join(GOLDEN, "biod_test_data")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def fixture_bytes(name):
    with open(os.path.join(BIOD_DATA, name), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def biod_data():
    return fixture_bytes
