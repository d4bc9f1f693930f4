import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def sample_excerpt():
    return np.load(os.path.join(GOLDEN, "sample_excerpt.npz"))


@pytest.fixture(scope="session")
def sample_full():
    return json.load(open(os.path.join(GOLDEN, "sample_full.json")))


@pytest.fixture(scope="session")
def synth_golden():
    return json.load(open(os.path.join(GOLDEN, "synth.json")))


def sha(a) -> str:
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
