import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement of the reference (oracle/), built on demand.  Checker only."""
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def lv():
    """The product package (lagrangianvoronoi.jl_b200, imported through the lvb200 alias)."""
    import lvb200
    return lvb200


def make_points(kind: str, n_side: int, seed: int = 0):
    """Seeded generator sets shared by the oracle and GPU parity tests.  Returns (xy, dr, bmin, bmax)."""
    from lvb200 import synthetic
    rng = np.random.default_rng(seed)
    dr = 1.0 / n_side
    if kind == "jitter":
        return synthetic.jittered_lattice(n_side, seed), dr, (0.0, 0.0), (1.0, 1.0)
    if kind == "poisson":
        return rng.random((n_side * n_side, 2)), dr, (0.0, 0.0), (1.0, 1.0)
    if kind == "rect2x1":  # non-square domain with an offset origin
        xy = rng.random((2 * n_side * n_side, 2)) * np.array([2.0, 1.0]) + np.array([-0.5, 0.25])
        return xy, dr, (-0.5, 0.25), (1.5, 1.25)
    if kind == "lattice":  # exact square lattice: four co-circular generators everywhere (degenerate)
        g = (np.arange(n_side) + 0.5) * dr
        X, Y = np.meshgrid(g, g, indexing="ij")
        return np.stack([X.ravel(), Y.ravel()], 1), dr, (0.0, 0.0), (1.0, 1.0)
    raise ValueError(kind)
