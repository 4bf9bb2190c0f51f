"""Device-resident time stepping: the callers either side of the hot path.

Same names and argument meaning as the reference's explicit operators (Julia's ``f!`` becomes ``f``); the
polygon fields live in HBM in label order between calls, so a canonical ``step!`` (examples/gresho.jl:100-114,
tests/taylorgreen.jl:61-71) runs without PCIe round trips:

    to_device(grid)                      upload grid.x and every field
    move(grid, dt)                       move!            move.jl:9-33        (remeshes)
    stiffened_eos(grid, gamma, P0)       stiffened_eos!   pressure.jl:64-70
    ideal_eos(grid, gamma, Pmin=0)       ideal_eos!       pressure.jl:49-55
    find_pressure_resident(solver, dt)   find_pressure!   pressure.jl:215-225
    pressure_step(grid, dt)              pressure_step!   pressure.jl:10-25
    gravity_step(grid, g, dt)            gravity_step!    pressure.jl:77-82
    find_D(grid)                         find_D!          diffusion.jl:8-19
    viscous_step(grid, dt, artificial_viscosity=True)     diffusion.jl:39-53
    find_dv(grid, dt, alpha=1.0)         find_dv!         relaxation.jl:10-25
    relaxation_step(grid, dt, rusanov=True)               relaxation.jl:36-73 (remeshes)
    from_device(grid)                    download x and every field (e.g. before export / postproc!)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import check, ptr
from .host import PressureSolver, VoronoiGrid, _FIELDS, _host_empty, _wall_velocities

_STATE_FIELDS = ["v", "dv", "momentum", "rho", "e", "P", "c2", "mass", "energy", "quality", "mu", "phase"]


def state_set(grid: VoronoiGrid, name: str, arr) -> None:
    a = np.ascontiguousarray(arr, dtype=np.float64)
    n = grid.n
    check(grid._L.lv_state_set(grid._h, name.encode(), ptr(a), n), grid._h)


def state_get(grid: VoronoiGrid, name: str, out=None) -> np.ndarray:
    nc = {"x": 2, "D": 4, "phase": 1}.get(name, _FIELDS.get(name, 1))
    n = grid.n
    if out is None:
        out = _host_empty((n, nc) if nc > 1 else (n,), np.float64)
    check(grid._L.lv_state_get(grid._h, name.encode(), ptr(out)), grid._h)
    return out


def to_device(grid: VoronoiGrid, remesh: bool = True) -> None:
    """Upload the generators and every polygon field; optionally build the mesh of the resident positions."""
    state_set(grid, "x", grid.x)
    for name in _STATE_FIELDS:
        state_set(grid, name, getattr(grid, name))
    if remesh:
        check(grid._L.lv_state_remesh(grid._h), grid._h)
    grid._resident = True


def from_device(grid: VoronoiGrid, mesh: bool = False) -> None:
    """Download positions and fields into the grid's host arrays (and, on request, the edge view)."""
    state_get(grid, "x", grid.x)
    for name in _STATE_FIELDS:
        state_get(grid, name, getattr(grid, name))
    if mesh:
        grid.rowptr, e, grid._area, grid._centroid = grid.mesh_download(grid.n)
        grid.edges = e


def remesh_resident(grid: VoronoiGrid) -> None:
    check(grid._L.lv_state_remesh(grid._h), grid._h)


def move(grid: VoronoiGrid, dt: float) -> None:
    check(grid._L.lv_step_move(grid._h, float(dt)), grid._h)


def stiffened_eos(grid: VoronoiGrid, gamma: float = 1.4, P0: float = 0.0) -> None:
    check(grid._L.lv_step_eos(grid._h, float(gamma), float(P0), 1), grid._h)


def ideal_eos(grid: VoronoiGrid, gamma: float = 1.4, Pmin: float = 0.0) -> None:
    check(grid._L.lv_step_eos(grid._h, float(gamma), float(Pmin), 0), grid._h)


def find_pressure_resident(solver: PressureSolver, dt: float, niter: int = 10, boundary_velocity=None) -> None:
    g = solver.grid
    iters = np.zeros(niter, np.int32)
    relres = np.zeros(niter) if solver.verbose else None
    vw, ve = _wall_velocities(g, boundary_velocity)
    check(g._L.lv_set_boundary_velocity(g._h, ptr(ve), 0 if ve is None else int(ve.shape[0])), g._h)
    kind = _capi.solver_kind(solver.solver)
    check(g._L.lv_step_find_pressure(g._h, float(dt), int(niter), solver.rtol, solver.atol, int(solver.itmax), kind, ptr(vw),
                                     iters.ctypes.data_as(C.POINTER(C.c_int32)),
                                     None if relres is None else relres.ctypes.data_as(C.POINTER(C.c_double))), g._h)
    solver.iters, solver.relres = iters, relres


def pressure_step(grid: VoronoiGrid, dt: float) -> None:
    check(grid._L.lv_step_pressure_step(grid._h, float(dt)), grid._h)


def gravity_step(grid: VoronoiGrid, g, dt: float) -> None:
    check(grid._L.lv_step_gravity(grid._h, float(g[0]), float(g[1]), float(dt)), grid._h)


def find_D(grid: VoronoiGrid) -> None:
    check(grid._L.lv_step_find_D(grid._h), grid._h)


def viscous_step(grid: VoronoiGrid, dt: float, artificial_viscosity: bool = True) -> None:
    check(grid._L.lv_step_viscous_step(grid._h, float(dt), int(artificial_viscosity)), grid._h)


def bdary_friction(grid: VoronoiGrid, dt: float, vwall=None, charfun=None) -> None:
    """bdary_friction!(grid, vDirichlet, dt; charfun)  diffusion.jl:64-80.

    ``vwall``: the reference's closure ``vDirichlet(m)`` (evaluated here at the midpoint of every boundary edge), or the four
    per-wall constants ``vwall[4][2]`` for UP, RIGHT, DOWN, LEFT (what examples/cavity.jl:41-44 amounts to).  ``charfun``: the
    reference's ``charfun(m)`` closure (e.g. top_and_bottom of examples/bubble.jl), or four per-wall flags; None = everywhere."""
    from .host import boundary_edges
    per_edge_v = callable(vwall)
    per_edge_c = callable(charfun)
    v_edge = on_edge = wall_on = None
    vw = np.zeros((4, 2))
    if per_edge_v or per_edge_c:
        mid, lab, _ = boundary_edges(grid)
        if per_edge_v:
            v_edge = np.ascontiguousarray([np.asarray(vwall(m), np.float64).reshape(2) for m in mid]).reshape(-1, 2)
        if per_edge_c:
            on_edge = np.ascontiguousarray([1 if charfun(m) else 0 for m in mid], dtype=np.uint8)
    if vwall is not None and not per_edge_v:
        vw = np.ascontiguousarray(vwall, dtype=np.float64)
        if vw.shape != (4, 2):
            raise ValueError("vwall must have shape (4, 2)")
    if charfun is not None and not per_edge_c:
        wall_on = np.ascontiguousarray(charfun, dtype=np.uint8).reshape(4)
    n_edge = 0 if (v_edge is None and on_edge is None) else int(len(v_edge) if v_edge is not None else len(on_edge))
    check(grid._L.lv_step_bdary_friction_ex(grid._h, float(dt), ptr(vw), ptr(wall_on), ptr(v_edge), ptr(on_edge), n_edge), grid._h)


def find_dv(grid: VoronoiGrid, dt: float, alpha: float = 1.0) -> None:
    check(grid._L.lv_step_find_dv(grid._h, float(dt), float(alpha)), grid._h)


def relaxation_step(grid: VoronoiGrid, dt: float, rusanov: bool = True) -> None:
    check(grid._L.lv_step_relaxation_step(grid._h, float(dt), int(rusanov)), grid._h)


def multiphase_projection(grid: VoronoiGrid, quality_threshold: float = 0.25, rtol: float = 1e-4, atol: float = 1e-4, itmax: int = 200):
    """multiphase_projection!(solver)  relaxation.jl:179-206 (MultiphaseSolver defaults).  Returns (iterations, solved)."""
    it, ok = C.c_int32(), C.c_int32()
    check(grid._L.lv_step_multiphase_projection(grid._h, float(quality_threshold), float(rtol), float(atol), int(itmax),
                                                C.byref(it), C.byref(ok)), grid._h)
    if not ok.value:
        import warnings
        warnings.warn("multiphase projector did not converge within tolerance")  # relaxation.jl:184-187
    return it.value, bool(ok.value)


def multiphase_apply(grid: VoronoiGrid, x):
    """(A x, b): mul!(res, A::MultiphaseProjector, x) (relaxation.jl:91-123) and the right-hand side refresh! builds from
    the resident dv (relaxation.jl:162-177), both on the device."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y, b = np.zeros_like(x), np.zeros_like(x)
    check(grid._L.lv_step_multiphase_apply(grid._h, ptr(x), ptr(y), ptr(b)), grid._h)
    return y, b
