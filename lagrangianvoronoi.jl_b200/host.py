"""Host-side mirror of the reference's Julia API for the mesh-and-pressure path.

Same names, argument meaning and error behaviour as the reference (Julia's ``remesh!`` /
``find_pressure!`` / ``mul!`` become ``remesh`` / ``find_pressure`` / ``mul``):

* ``VoronoiGrid(boundary_rect, dr; h, r_max, xperiodic, yperiodic)``  voronoigrid.jl:26-49
* ``remesh(grid)``                                                    voronoigrid.jl:89-108
* ``PressureSolver(grid; verbose)``                                   pressure.jl:150-158
* ``find_pressure(solver, dt, niter=10; boundary_velocity)``          pressure.jl:215-225
* ``mul(y, A, x)``                                                    pressure.jl:119-130

The reference stores polygons as an array of heap structs; here the polygons are a
structure of numpy arrays in label order (``grid.x``, ``grid.v``, ``grid.P`` ...) and
``grid.edges`` / ``grid.rowptr`` are the flat view of every ``p.edges`` (40-byte ``Edge``
records, polygon ``i`` owns ``edges[rowptr[i]:rowptr[i+1]]``), filled by one device->host
copy.  All computation happens in liblvb200.so; nothing here has a CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _capi
from ._capi import EDGE_DTYPE, LV_SOLVER_CG, LV_SOLVER_MINRES, check, load_library, ptr

BDARY_UP, BDARY_RIGHT, BDARY_DOWN, BDARY_LEFT = -1, -2, -3, -4  # polygon.jl:4-7

# per-polygon fields of @Euler_vars (celldefs.jl:7-27): name -> components
_FIELDS = {"rho": 1, "v": 2, "e": 1, "P": 1, "c2": 1, "dv": 2, "mass": 1, "momentum": 2, "energy": 1, "quality": 1,
           "mu": 1, "phase": 1}


def _host_empty(shape, dtype):
    """Pinned host array when CUDA is up (async copies at full PCIe rate), plain numpy otherwise."""
    try:
        import torch
        if torch.cuda.is_available():
            t = torch.empty(int(np.prod(shape)) * np.dtype(dtype).itemsize, dtype=torch.uint8, pin_memory=True)
            a = t.numpy().view(dtype).reshape(shape)
            return a
    except Exception:  # pragma: no cover - torch missing or no driver
        pass
    return np.empty(shape, dtype)


@dataclass
class Rectangle:
    """geometry.jl:94-98"""
    xmin: tuple
    xmax: tuple

    @staticmethod
    def from_lims(xlims, ylims):  # geometry.jl:107-111
        return Rectangle((float(xlims[0]), float(ylims[0])), (float(xlims[1]), float(ylims[1])))


class VoronoiGrid:
    """voronoigrid.jl:14-50.  Also owns the device context of liblvb200."""

    def __init__(self, boundary_rect: Rectangle, dr: float, h: float | None = None, r_max: float | None = None,
                 xperiodic: bool = False, yperiodic: bool = False, device: int = 0):
        L = load_library()
        self.dr = float(dr)
        self.h = 2.0 * self.dr if h is None else float(h)
        r_max = 10.0 * self.dr if r_max is None else float(r_max)
        self.rr_max = r_max * r_max
        self.xperiodic, self.yperiodic = bool(xperiodic), bool(yperiodic)
        self.boundary_rect = boundary_rect
        bmin = np.asarray(boundary_rect.xmin, np.float64)
        bmax = np.asarray(boundary_rect.xmax, np.float64)
        # voronoigrid.jl:29-33
        self.cropping_rect = Rectangle(tuple(bmin - r_max * np.array([xperiodic, yperiodic], np.float64)),
                                       tuple(bmax + r_max * np.array([xperiodic, yperiodic], np.float64)))
        self.xperiod = float(bmax[0] - bmin[0])
        self.yperiod = float(bmax[1] - bmin[1])
        if self.h <= 0.0:
            raise ValueError("h must be positive")  # neighborlist.jl:19-21
        desc = _capi.GridDesc(self.dr, self.h, r_max, int(self.xperiodic), int(self.yperiodic),
                              (C.c_double * 2)(*bmin), (C.c_double * 2)(*bmax))
        self._h = C.c_void_p()
        check(L.lv_create(C.byref(desc), int(device), C.byref(self._h)), None)
        self._L = L
        self.device = int(device)
        # polygons (structure of arrays, label order)
        self.x = np.zeros((0, 2))
        for name, nc in _FIELDS.items():
            setattr(self, name, np.zeros((0, nc) if nc > 1 else 0))
        self.rowptr = None
        self.edges = None
        self._area = None
        self._centroid = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._L.lv_destroy(h)
            except Exception:
                pass
            self._h = None

    # -- polygons ------------------------------------------------------------------------------
    @property
    def n(self) -> int:
        return int(self.x.shape[0])

    def set_points(self, xy) -> None:
        """push!(grid.polygons, T(x=x)) for every row of xy; fields take the celldefs.jl defaults."""
        xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        n = xy.shape[0]
        self.x = _host_empty((n, 2), np.float64)
        self.x[...] = xy
        for name, nc in _FIELDS.items():
            a = _host_empty((n, nc) if nc > 1 else (n,), np.float64)
            a[...] = 0.0
            setattr(self, name, a)
        self.rowptr = self.edges = self._area = self._centroid = None

    def set_rects(self, boundary_rect: Rectangle, cropping_rect: Rectangle) -> None:
        """examples/piston.jl:43-47 mutates both rectangles at run time."""
        self.boundary_rect, self.cropping_rect = boundary_rect, cropping_rect
        a = [np.asarray(v, np.float64) for v in (boundary_rect.xmin, boundary_rect.xmax, cropping_rect.xmin,
                                                   cropping_rect.xmax)]
        check(self._L.lv_set_rects(self._h, *[v.ctypes.data_as(C.POINTER(C.c_double)) for v in a]), self._h)

    def info(self) -> dict:
        n1, n2, npath = C.c_int64(), C.c_int64(), C.c_int64()
        origin = (C.c_double * 2)()
        check(self._L.lv_grid_info(self._h, C.byref(n1), C.byref(n2), origin, C.byref(npath)), self._h)
        return {"n1": n1.value, "n2": n2.value, "npath": npath.value, "origin": np.array(origin[:])}

    def magic_path(self):
        m = self.info()["npath"]
        i1, i2, rr = np.zeros(m, np.int64), np.zeros(m, np.int64), np.zeros(m)
        cnt = C.c_int64()
        ip = C.POINTER(C.c_int64)
        check(self._L.lv_magic_path(self._h, m, i1.ctypes.data_as(ip), i2.ctypes.data_as(ip),
                                    rr.ctypes.data_as(C.POINTER(C.c_double)), C.byref(cnt)), self._h)
        return i1, i2, rr

    # -- instrumentation ---------------------------------------------------------------------------
    def prof_enable(self, on: bool = True):
        check(self._L.lv_prof_enable(self._h, int(on)), self._h)

    def prof_reset(self):
        check(self._L.lv_prof_reset(self._h), self._h)

    def prof_get(self, slot: str):
        ms, cnt = C.c_double(), C.c_int64()
        check(self._L.lv_prof_get(self._h, _capi.PROF_SLOTS[slot], C.byref(ms), C.byref(cnt)), self._h)
        return ms.value, cnt.value

    def clip_info(self):
        """(level, anomalies): which clipping kernel produced the current mesh (see lv_capi.h)."""
        lvl, an = C.c_int32(), C.c_int64()
        check(self._L.lv_clip_info(self._h, C.byref(lvl), C.byref(an)), self._h)
        return lvl.value, an.value

    def launch_count(self) -> int:
        return int(self._L.lv_launch_count(self._h))

    def device_bytes(self) -> int:
        return int(self._L.lv_device_bytes(self._h))

    def sync(self):
        check(self._L.lv_sync(self._h), self._h)

    def set_stream(self, stream_ptr: int | None):
        """Run on the given cudaStream_t (0 = the legacy default stream torch uses); None = the handle's own stream."""
        if stream_ptr is None:
            check(self._L.lv_set_stream(self._h, None, 1), self._h)
        else:
            check(self._L.lv_set_stream(self._h, C.c_void_p(int(stream_ptr)), 0), self._h)

    # -- device-resident entry points (inputs already in HBM; used by bench.py `value`) ---------------
    def remesh_dev(self, xy_dev, n: int | None = None) -> None:
        """remesh! on positions that already live on the GPU (torch tensor or raw device pointer)."""
        if n is None:
            n = int(xy_dev.shape[0])
        check(self._L.lv_remesh_dev(self._h, int(n), ptr(xy_dev)), self._h)
        self._dev_n = int(n)

    def mesh_nnz(self) -> int:
        nnz = C.c_int64()
        check(self._L.lv_mesh_nnz(self._h, C.byref(nnz)), self._h)
        return nnz.value

    def mesh_download(self, n: int | None = None, edges: bool = True):
        n = self.n if n is None else n
        nnz = self.mesh_nnz()
        rowptr = _host_empty((n + 1,), np.int64)
        e = _host_empty((nnz,), EDGE_DTYPE) if edges else None
        ar = _host_empty((n,), np.float64)
        ce = _host_empty((n, 2), np.float64)
        check(self._L.lv_mesh_download(self._h, ptr(rowptr), ptr(e), nnz, ptr(ar), ptr(ce)), self._h)
        if getattr(self, "_lazy_edges", 0) == 3:  # pipelined mode: the download runs in the background
            check(self._L.lv_mesh_wait(self._h), self._h)
        return rowptr, e, ar, ce


def wait_edges(grid: VoronoiGrid) -> None:
    """Block until the edge view of the last lazy remesh has landed in ``grid.edges``.  In the pipelined mode
    (``lazy="pipeline"``) this is also where a deferred remesh reports its errors and where ``grid.edges`` gets its length."""
    check(grid._L.lv_mesh_wait(grid._h), grid._h)
    if getattr(grid, "_lazy_edges", 0) == 3 and getattr(grid, "_edge_buf", None) is not None and grid.n:
        grid.edges = grid._edge_buf[: grid.mesh_nnz()]


def remesh(grid: VoronoiGrid, edges: bool = True, lazy: bool = False) -> None:
    """remesh!(grid)  voronoigrid.jl:89-108.

    ``lazy=True`` returns once ``rowptr``, areas and centroids are back and lets the edge records (40 B each, the bulk
    of the traffic) arrive in the background; call ``wait_edges(grid)`` before reading ``grid.edges``.  ``lazy="all"``
    sends rowptr, areas and centroids in the background as well: nothing may be read before ``wait_edges(grid)``.
    ``lazy="pipeline"`` additionally returns while the clipping kernel is still queued (the next call's uploads overlap
    it; errors of this remesh surface at the next call or at ``wait_edges``) and moves the mesh over PCIe in a 20 B/edge
    wire format that host threads of the library expand into the 40-byte records (lv_pipeline.cu).

    Gathers ``grid.x``, runs the cell-list build and the clipping kernel on the GPU and leaves
    ``grid.rowptr`` / ``grid.edges`` (the flat ``p.edges`` view), areas and centroids on the host.
    Raises LvError("The Voronoi Mesh has been destroyed.") like voronoigrid.jl:63-65.
    """
    n = grid.n
    L = grid._L
    if grid.rowptr is None or grid.rowptr.shape[0] != n + 1:
        grid.rowptr = _host_empty((n + 1,), np.int64)
        grid._area = _host_empty((n,), np.float64)
        grid._centroid = _host_empty((n, 2), np.float64)
    nnz = C.c_int64()
    if not edges:
        if getattr(grid, "_lazy_edges", 0) == 3:  # rowptr / areas / centroids are wanted on return: leave the pipelined mode
            check(L.lv_set_async_edges(grid._h, 0), grid._h)
            grid._lazy_edges = 0
        check(L.lv_remesh(grid._h, n, ptr(grid.x), ptr(grid.rowptr), None, 0, C.byref(nnz), ptr(grid._area),
                          ptr(grid._centroid)), grid._h)
        grid.edges = None
        return
    mode = 3 if lazy == "pipeline" else (2 if lazy == "all" else int(bool(lazy)))
    if mode != getattr(grid, "_lazy_edges", 0):
        check(L.lv_set_async_edges(grid._h, mode), grid._h)
        grid._lazy_edges = mode
    buf = getattr(grid, "_edge_buf", None)
    if buf is None or buf.shape[0] < 6 * n + 64:  # grow-only pinned buffer: page-locking GBs per call would dominate
        wait_edges(grid)  # a background copy of the previous lazy remesh may still be writing into the old buffer
        grid._edge_buf = _host_empty((7 * n + 64,), EDGE_DTYPE)
    st = L.lv_remesh(grid._h, n, ptr(grid.x), ptr(grid.rowptr), ptr(grid._edge_buf), grid._edge_buf.shape[0],
                     C.byref(nnz), ptr(grid._area), ptr(grid._centroid))
    if st == _capi.LV_ECAPACITY and nnz.value > grid._edge_buf.shape[0]:  # pragma: no cover - 7n is generous
        wait_edges(grid)
        grid._edge_buf = _host_empty((nnz.value,), EDGE_DTYPE)
        st = L.lv_mesh_download(grid._h, ptr(grid.rowptr), ptr(grid._edge_buf), nnz.value, ptr(grid._area),
                                ptr(grid._centroid))
    check(st, grid._h)
    grid.edges = grid._edge_buf[: nnz.value] if nnz.value >= 0 else None  # pipelined: known after wait_edges


def area(grid: VoronoiGrid) -> np.ndarray:
    """area(p) for every polygon (polygon.jl:114-122), computed by the clipping kernel."""
    if grid._area is None:
        raise RuntimeError("remesh(grid) first")
    return grid._area


def centroid(grid: VoronoiGrid) -> np.ndarray:
    """centroid(p) for every polygon (polygon.jl:210-219)."""
    if grid._centroid is None:
        raise RuntimeError("remesh(grid) first")
    return grid._centroid


def neighbors_csr(grid: VoronoiGrid):
    """(rowptr, labels) of neighbors(p, grid) (iterators.jl:23-33): edges with label > 0, 1-based."""
    lab = grid.edges["label"]
    keep = lab > 0
    cnt = np.add.reduceat(keep.astype(np.int64), grid.rowptr[:-1]) if grid.n else np.zeros(0, np.int64)
    cnt[np.diff(grid.rowptr) == 0] = 0
    rp = np.concatenate([[0], np.cumsum(cnt)])
    return rp, lab[keep]


class PressureSolver:
    """PressureSolver(grid; verbose)  pressure.jl:142-159: device workspace for the current n."""

    def __init__(self, grid: VoronoiGrid, verbose: bool = False, solver: str = "cg", rtol: float = 1e-6,
                 atol: float = 1e-6, itmax: int = 1000):
        self.grid = grid
        self.verbose = verbose
        self.solver = solver
        self.rtol, self.atol, self.itmax = rtol, atol, itmax  # pressure.jl:219 defaults
        self.iters = None
        self.relres = None
        check(grid._L.lv_pressure_create(grid._h), grid._h)

    # -- device-resident entry points -----------------------------------------------------------------
    def upload_fields(self, mass=None, rho=None, c2=None, P=None, v=None, device: bool = False):
        g = self.grid
        fn = g._L.lv_fields_upload_dev if device else g._L.lv_fields_upload
        check(fn(g._h, ptr(mass), ptr(rho), ptr(c2), ptr(P), ptr(v)), g._h)

    def assemble(self, dt: float):
        g = self.grid
        check(g._L.lv_pressure_assemble(g._h, float(dt)), g._h)

    def operator(self):
        """(rowptr, col (1-based), w, diag) = A.neighbors / A.lr_ratios / A.diagonal  pressure.jl:89-93"""
        g = self.grid
        n = g.n if g.n else getattr(g, "_dev_n", 0)
        nnz = g.mesh_nnz()
        rowptr = np.zeros(n + 1, np.int64)
        col = np.zeros(nnz, np.int64)
        w = np.zeros(nnz)
        diag = np.zeros(n)
        check(g._L.lv_pressure_operator(g._h, ptr(rowptr), ptr(col), ptr(w), nnz, ptr(diag)), g._h)
        m = int(rowptr[-1])
        return rowptr, col[:m], w[:m], diag

    def set_boundary_velocity(self, vbc_edge=None):
        """Per-boundary-edge wall velocities (order of ``boundary_edges``) for the following right-hand sides; None clears."""
        g = self.grid
        ve = None if vbc_edge is None else np.ascontiguousarray(vbc_edge, np.float64).reshape(-1, 2)
        check(g._L.lv_set_boundary_velocity(g._h, ptr(ve), 0 if ve is None else int(ve.shape[0])), g._h)

    def rhs(self, dt: float, gp_step: bool = False, vbc_wall=None, vbc_edge=None):
        g = self.grid
        n = g.n if g.n else getattr(g, "_dev_n", 0)
        b, GP = np.zeros(n), np.zeros((n, 2))
        vw = None if vbc_wall is None else np.ascontiguousarray(vbc_wall, np.float64).reshape(4, 2)
        self.set_boundary_velocity(vbc_edge)
        check(g._L.lv_pressure_rhs(g._h, float(dt), int(gp_step), ptr(vw), ptr(b), ptr(GP)), g._h)
        return b, GP

    def solve(self, b, x0, rtol=1e-10, atol=0.0, itmax=100000, solver=None):
        g = self.grid
        b = np.ascontiguousarray(b, np.float64)
        x = np.array(x0, np.float64, copy=True)
        it, rr = C.c_int32(), C.c_double()
        kind = _capi.solver_kind(solver or self.solver)
        check(g._L.lv_pressure_solve(g._h, kind, ptr(b), ptr(x), rtol, atol, int(itmax), C.byref(it), C.byref(rr)), g._h)
        return x, it.value, rr.value

    def find_pressure_dev(self, dt, niter=10, vbc_wall=None, want_relres=False):
        g = self.grid
        iters = np.zeros(niter, np.int32)
        relres = np.zeros(niter) if want_relres else None
        vw = None if vbc_wall is None else np.ascontiguousarray(vbc_wall, np.float64).reshape(4, 2)
        kind = _capi.solver_kind(self.solver)
        check(g._L.lv_find_pressure_dev(g._h, float(dt), int(niter), self.rtol, self.atol, int(self.itmax), kind, ptr(vw),
                                        iters.ctypes.data_as(C.POINTER(C.c_int32)),
                                        None if relres is None else relres.ctypes.data_as(C.POINTER(C.c_double))), g._h)
        self.iters, self.relres = iters, relres
        return iters, relres

    def download_P(self, out=None):
        g = self.grid
        n = g.n if g.n else getattr(g, "_dev_n", 0)
        out = _host_empty((n,), np.float64) if out is None else out
        check(g._L.lv_pressure_download(g._h, ptr(out)), g._h)
        return out


def boundary_edges(grid: VoronoiGrid):
    """(midpoint[nb, 2], label[nb], polygon[nb]) of every boundary edge -- boundaries(p) (iterators.jl:50-57) over the whole
    mesh, polygon by polygon in label order, edges in p.edges order.  ``polygon`` is 1-based like the reference's indices."""
    L = grid._L
    cnt = C.c_int64()
    check(L.lv_boundary_edges(grid._h, C.byref(cnt), None, None, None, 0), grid._h)
    nb = cnt.value
    mid, lab, pol = np.zeros((nb, 2)), np.zeros(nb, np.int64), np.zeros(nb, np.int64)
    if nb:
        check(L.lv_boundary_edges(grid._h, C.byref(cnt), ptr(mid), ptr(lab), ptr(pol), nb), grid._h)
    return mid, lab, pol


def _wall_velocities(grid: VoronoiGrid, boundary_velocity):
    """The reference passes a closure boundary_velocity(midpoint(e), e.label) (pressure.jl:182).  No callback crosses the
    C ABI: the closure is evaluated HERE at the midpoint of every boundary edge.  Returns (vbc_wall, vbc_edge): when the
    values are constant along every wall (what all the reference's examples do, e.g. examples/piston.jl:122-127) the four
    per-wall constants and None -- the fast path; otherwise None and the per-edge array.  An explicit (4, 2) array of wall
    constants or an (nb, 2) array of per-edge values is accepted as well."""
    if boundary_velocity is None:
        return None, None
    if not callable(boundary_velocity):
        a = np.ascontiguousarray(boundary_velocity, np.float64)
        if a.shape == (4, 2) or a.size == 8:
            return a.reshape(4, 2), None
        return None, a.reshape(-1, 2)
    mid, lab, _ = boundary_edges(grid)
    if len(lab) == 0:
        return None, None
    vals = np.array([np.asarray(boundary_velocity(m, int(l)), np.float64).reshape(2) for m, l in zip(mid, lab)])
    wall = np.zeros((4, 2))
    constant = True
    for code in (BDARY_UP, BDARY_RIGHT, BDARY_DOWN, BDARY_LEFT):
        sel = lab == code
        if sel.any():
            wall[-code - 1] = vals[sel][0]
            constant &= bool((vals[sel] == vals[sel][0]).all())
    if constant and np.isin(lab, (-1, -2, -3, -4)).all():
        return wall, None
    return None, np.ascontiguousarray(vals)


def find_pressure(solver: PressureSolver, dt: float, niter: int = 10, boundary_velocity=None) -> None:
    """find_pressure!(solver, dt, niter; boundary_velocity)  pressure.jl:215-225.

    Gathers mass, rho, c2, P, v from the grid, runs operator assembly + ``niter`` x (RHS + Krylov
    solve) on the GPU and scatters the pressure back into ``grid.P``."""
    g = solver.grid
    n = g.n
    iters = np.zeros(niter, np.int32)
    relres = np.zeros(niter) if solver.verbose else None
    vw, ve = _wall_velocities(g, boundary_velocity)
    kind = _capi.solver_kind(solver.solver)
    P_out = g.P
    check(g._L.lv_find_pressure(g._h, float(dt), int(niter), solver.rtol, solver.atol, int(solver.itmax), kind,
                                ptr(g.mass), ptr(g.rho), ptr(g.c2), ptr(g.P), ptr(g.v), ptr(vw), ptr(ve),
                                0 if ve is None else int(ve.shape[0]), ptr(P_out),
                                iters.ctypes.data_as(C.POINTER(C.c_int32)),
                                None if relres is None else relres.ctypes.data_as(C.POINTER(C.c_double))), g._h)
    solver.iters, solver.relres = iters, relres
    if solver.verbose:
        print(f"find_pressure: n={n} iterations={iters.tolist()} relres={relres.tolist()}")


def mul(y: np.ndarray, A: PressureSolver, x: np.ndarray) -> np.ndarray:
    """mul!(y, A, x)  pressure.jl:119-130 on the operator assembled by the last find_pressure/assemble."""
    g = A.grid
    x = np.ascontiguousarray(x, np.float64)
    check(g._L.lv_pressure_matvec(g._h, ptr(x), ptr(y)), g._h)
    return y
