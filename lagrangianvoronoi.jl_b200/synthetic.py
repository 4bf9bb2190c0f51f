"""Deterministic synthetic inputs of the benchmark.

Unit square, periodic in x and y, M x M generators on a lattice jittered by +-0.4 dr with the
splitmix64 finaliser, so that the CPU checker, the CUDA path and (if ever available) Julia consume
bit-identical positions without porting an RNG.
"""
from __future__ import annotations

import numpy as np

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)


def mix64(z: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    z = z.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return z


def uniform(seed: int, counter: np.ndarray) -> np.ndarray:
    """U(c) = (mix64(seed*0x9E3779B97F4A7C15 + c) >> 11) * 2^-53"""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) * _GOLDEN + counter.astype(np.uint64)
    return (mix64(z) >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def jittered_lattice(M: int, seed: int = 0, jitter: float = 0.8, rows: tuple | None = None, My: int | None = None,
                     return_index: bool = False):
    """Generator k = i*My + j at ((i+0.5+J(u1-0.5))dr, (j+0.5+J(u2-0.5))dr), dr = 1/M, i < M, j < My (= M).

    ``rows=(j0, j1)`` returns only the generators with lattice index j in [j0, j1) -- the y-strip a rank owns in
    the multi-GPU decomposition -- in the same global order.  ``My`` stretches the box to [0,1] x [0, My/M]
    (weak scaling: one unit square per GPU).  ``return_index`` also returns the 0-based global index k."""
    dr = 1.0 / M
    My = M if My is None else My
    i = np.arange(M, dtype=np.int64)
    j = np.arange(My, dtype=np.int64) if rows is None else np.arange(rows[0], rows[1], dtype=np.int64)
    I, J = np.meshgrid(i, j, indexing="ij")
    k = (I * My + J).reshape(-1)
    u1 = uniform(seed, 2 * k)
    u2 = uniform(seed, 2 * k + 1)
    x = (I.reshape(-1) + 0.5 + jitter * (u1 - 0.5)) * dr
    y = (J.reshape(-1) + 0.5 + jitter * (u2 - 0.5)) * dr
    xy = np.stack([x, y], axis=1)
    return (xy, k) if return_index else xy


def taylor_green_fields(xy: np.ndarray, t: float = 0.0, Re: float = 400.0):
    """Analytic Taylor-Green velocity and pressure (tests/taylorgreen.jl:35-43)."""
    vmax = np.exp(-8.0 * np.pi ** 2 * t / Re)
    u0 = np.cos(2 * np.pi * xy[:, 0]) * np.sin(2 * np.pi * xy[:, 1])
    v0 = -np.sin(2 * np.pi * xy[:, 0]) * np.cos(2 * np.pi * xy[:, 1])
    v = vmax * np.stack([u0, v0], axis=1)
    P = 0.5 * vmax ** 2 * (np.sin(2 * np.pi * xy[:, 0]) ** 2 + np.sin(2 * np.pi * xy[:, 1]) ** 2 - 1.0)
    return v, P
