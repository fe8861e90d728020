import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def oracle():
    from oracle import c4oracle
    c4oracle.build()
    return c4oracle


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def random_positions(seed, n, max_plies=34):
    """SURVEY.md 8(d) config 2 sampling (same as tests/golden/generate_goldens.py), on the oracle's board."""
    import random
    from oracle import c4oracle as o
    rng = random.Random(seed)
    out = []
    while len(out) < n:
        plies = rng.randint(0, max_plies)
        c0 = c1 = 0
        ok = True
        for _ in range(plies):
            m = o.legal_mask(c0, c1)
            c0, c1, res = o.drop(c0, c1, rng.choice([c for c in range(7) if m >> c & 1]))
            if res != -1:
                ok = False
                break
        if ok:
            out.append((c0, c1))
    a = np.array(out, dtype=np.uint64).reshape(-1, 2)
    return a[:, 0].copy(), a[:, 1].copy()
