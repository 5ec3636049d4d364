import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
for p in (os.path.join(ROOT, "mcmurchie-davidson_b200"), ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu():
    try:
        from mmd._b200 import lib
        lib.require_gpu()
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        path = os.path.join(GOLDEN, name)
        if name.endswith(".json"):
            with open(path) as f:
                return json.load(f)
        return np.load(path)
    return load


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


def unique_pack(T):
    """Canonical unique integrals in the reference's doERIs loop order (i>=j, k>=l, ij>=kl)."""
    N = T.shape[0]
    i, j = np.tril_indices(N)
    ij = i * (i + 1) // 2 + j
    I, K = np.meshgrid(np.arange(len(ij)), np.arange(len(ij)), indexing="ij")
    keep = ij[I] >= ij[K]
    # order: i, j, k, l nested loops with k over 0..N-1, l<=k  == row-major over (ij index, kl index)
    a, b = I[keep], K[keep]
    return T[i[a], j[a], i[b], j[b]]
