import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU check, excluded from the default CPU suite")


@pytest.fixture(scope="session")
def golden_cases():
    import json

    with open(os.path.join(ROOT, "tests", "golden", "reference_cases.json")) as fh:
        return json.load(fh)
