"""Seeding strategies of the reference (populate.jl): each pushes generators, calls ``remesh!`` and then the
initial condition.  ``ic`` receives the grid (structure of arrays) instead of one polygon at a time:
``ic(grid)`` may fill ``grid.v``, ``grid.rho``, ``grid.mass`` ... using ``area(grid)``.

populate_lloyd! (populate.jl:132-145) runs its 100 x (remesh! + centroid move) loop on the device
(``lv_step_lloyd``), the other strategies only generate points on the host (one-off work).
"""
from __future__ import annotations

import numpy as np

from ._capi import check
from .host import VoronoiGrid, remesh


def _finish(grid: VoronoiGrid, pts, charfun, ic, edges=True):
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    (x0, y0), (x1, y1) = grid.boundary_rect.xmin, grid.boundary_rect.xmax
    keep = (pts[:, 0] >= x0) & (pts[:, 0] <= x1) & (pts[:, 1] >= y0) & (pts[:, 1] <= y1)  # isinside  geometry.jl:127-129
    if charfun is not None:
        keep &= np.asarray(charfun(pts), dtype=bool)
    grid.set_points(pts[keep])
    remesh(grid, edges=edges)
    if ic is not None:
        ic(grid)


def _julia_linrange(a: float, b: float, n: int) -> np.ndarray:
    """range(a, b, n): n points, both ends included.  Julia evaluates ranges in twice precision, i.e. every element is the
    correctly rounded value of a + i (b - a)/(n - 1); 80-bit long doubles reproduce that."""
    L = np.longdouble
    t = np.arange(n, dtype=L) / L(max(n - 1, 1))
    return (L(a) + t * (L(b) - L(a))).astype(np.float64)


def populate_rect(grid: VoronoiGrid, charfun=None, ic=None):
    """populate_rect!  populate.jl:46-66"""
    (x0, y0), (x1, y1) = grid.boundary_rect.xmin, grid.boundary_rect.xmax
    N = int(round((x1 - x0) / grid.dr)); M = int(round((y1 - y0) / grid.dr))
    X, Y = np.meshgrid(_julia_linrange(x0, x1, N), _julia_linrange(y0, y1, M), indexing="ij")
    _finish(grid, np.stack([X.ravel() + 0.5 * grid.dr, Y.ravel() + 0.5 * grid.dr], 1), charfun, ic)


def populate_hex(grid: VoronoiGrid, charfun=None, ic=None):
    """populate_hex!  populate.jl:149-174"""
    a = (4 / 3) ** 0.25 * grid.dr
    b = (3 / 4) ** 0.25 * grid.dr
    (x0, y0), (x1, y1) = grid.boundary_rect.xmin, grid.boundary_rect.xmax
    i = np.arange(int(np.floor(x0 / a)) - 1, int(np.ceil(x1 / a)) + 1)
    j = np.arange(int(np.floor(y0 / b)), int(np.ceil(y1 / b)) + 1)
    I, J = np.meshgrid(i, j, indexing="ij")
    jm = np.fmod(J, 2)  # Julia's % keeps the sign of the dividend
    _finish(grid, np.stack([((I + jm / 2) * a).ravel(), (J * b).ravel()], 1), charfun, ic)


def populate_rand(grid: VoronoiGrid, charfun=None, ic=None, seed: int = 0, samples=None):
    """populate_rand!  populate.jl:76-95.  Julia's global RNG is replaced by a seeded numpy generator, or by the caller's own
    uniform samples ``samples[N, 2]`` (e.g. read from a file shared with other implementations)."""
    (x0, y0), (x1, y1) = grid.boundary_rect.xmin, grid.boundary_rect.xmax
    N = int(round(abs(x1 - x0) * abs(y1 - y0) / grid.dr ** 2))
    s = np.random.default_rng(seed).random((N, 2)) if samples is None else np.asarray(samples, np.float64).reshape(-1, 2)[:N]
    _finish(grid, np.stack([s[:, 0] * x1 + (1 - s[:, 0]) * x0, s[:, 1] * y1 + (1 - s[:, 1]) * y0], 1), charfun, ic)


def populate_circ(grid: VoronoiGrid, charfun=None, center=(0.0, 0.0), ic=None):
    """populate_circ!  populate.jl:18-35"""
    (x0, y0), (x1, y1) = grid.boundary_rect.xmin, grid.boundary_rect.xmax
    c = np.asarray(center, dtype=np.float64)
    r_max = max(np.hypot(px - c[0], py - c[1]) for px in (x0, x1) for py in (y0, y1))
    pts = []
    dr = grid.dr
    # Julia's range (0.5dr):dr:r_max has floor((r_max - 0.5dr)/dr) + 1 elements (the end point included when it is hit)
    nring = int(np.floor((r_max - 0.5 * dr) / dr * (1.0 + 4e-16))) + 1 if r_max >= 0.5 * dr else 0
    for r in (float(np.longdouble(0.5) * np.longdouble(dr) + np.longdouble(k) * np.longdouble(dr)) for k in range(nring)):  # twice precision
        k_max = int(round(2.0 * np.pi * r / grid.dr))
        th = 2.0 * np.pi * np.arange(1, k_max + 1) / max(k_max, 1)
        pts.append(np.stack([c[0] + r * np.cos(th), c[1] + r * np.sin(th)], 1))
    _finish(grid, np.concatenate(pts) if pts else np.zeros((0, 2)), charfun, ic)


def populate_vogel(grid: VoronoiGrid, charfun=None, center=(0.0, 0.0), ic=None):
    """populate_vogel!  populate.jl:105-121"""
    (x0, y0), (x1, y1) = grid.boundary_rect.xmin, grid.boundary_rect.xmax
    c = np.asarray(center, dtype=np.float64)
    r_max = max(np.hypot(px - c[0], py - c[1]) for px in (x0, x1) for py in (y0, y1))
    N = int(round(np.pi * r_max * r_max / grid.dr ** 2))
    i = np.arange(1, N + 1)
    r = r_max * np.sqrt(i / N)
    th = 2.39996322972865332 * i
    _finish(grid, np.stack([c[0] + r * np.cos(th), c[1] + r * np.sin(th)], 1), charfun, ic)


def populate_lloyd(grid: VoronoiGrid, charfun=None, niterations: int = 100, ic=None, seed: int = 0, samples=None):
    """populate_lloyd!  populate.jl:132-145: random seeding, then niterations x (remesh!; x = centroid) on the device."""
    from . import stepping
    populate_rand(grid, charfun=charfun, ic=None, seed=seed, samples=samples)
    stepping.state_set(grid, "x", grid.x)
    check(grid._L.lv_step_lloyd(grid._h, int(niterations)), grid._h)
    stepping.state_get(grid, "x", grid.x)
    grid.rowptr, grid.edges, grid._area, grid._centroid = grid.mesh_download(grid.n)
    if ic is not None:
        ic(grid)
