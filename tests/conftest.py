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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def frames():
    with open(os.path.join(GOLDEN, "intel_gfs_head.json")) as f:
        return json.load(f)["frames"]


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def reading(fr):
    return {"x": fr["x"], "y": fr["y"], "theta": fr["theta"], "range": fr["range"]}


def dense_counts(G, cells, visited, total):
    v = np.ones(G * G)
    t = 2 * np.ones(G * G)
    v[cells] = visited
    t[cells] = total
    return v.reshape(G, G), t.reshape(G, G)
