import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))  # `pyref` package (oracle/pyref) — tests only


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def goldens():
    import json
    g = os.path.join(ROOT, "tests", "golden")
    return {
        "ref": json.load(open(os.path.join(g, "reference_kats.json"))),
        "derived": json.load(open(os.path.join(g, "derived_vectors.json"))),
        "poseidon_constants": json.load(open(os.path.join(g, "poseidon_constants.json"))),
    }
