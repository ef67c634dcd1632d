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


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def board_from_hex(rows, cols, hex_rows=None):
    """float64 grid (1.0 = occupied) from the fixtures' per-row hex occupancy."""
    g = np.zeros((rows, cols))
    if hex_rows is not None:
        for r, h in enumerate(hex_rows):
            v = int(h, 16)
            for c in range(cols):
                if (v >> c) & 1:
                    g[r, c] = 1.0
    return g


def hex_rows_of(grid):
    out = []
    for row in np.asarray(grid):
        v = 0
        for c, x in enumerate(row):
            if x != 0:
                v |= 1 << c
        out.append(hex(v))
    return out


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]
    return get
