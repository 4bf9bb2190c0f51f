"""ctypes wrapper over oracle/liblvoracle.so -- the CPU restatement of the reference path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
Parity status: see oracle/lv_oracle.h ("parity unpinned" at the Krylov-iterate level).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblvoracle.so")

EDGE_DTYPE = np.dtype([("v1", "<f8", (2,)), ("v2", "<f8", (2,)), ("label", "<i8")])  # geometry.jl:82-87
assert EDGE_DTYPE.itemsize == 40


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc -O2 -ffp-contract=off -fopenmp)."""
    src = os.path.join(_HERE, "lv_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_void_p
        L.lvo_grid_create.restype = vp
        L.lvo_grid_create.argtypes = [dp, dp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int]
        L.lvo_grid_destroy.argtypes = [vp]
        L.lvo_set_rects.argtypes = [vp, dp, dp, dp, dp]
        L.lvo_grid_info.argtypes = [vp, ip, ip, ip, dp]
        L.lvo_magic_path.restype = C.c_int64
        L.lvo_magic_path.argtypes = [vp, C.c_int64, ip, ip, dp]
        L.lvo_set_points.argtypes = [vp, C.c_int64, dp]
        L.lvo_npolygons.restype = C.c_int64
        L.lvo_npolygons.argtypes = [vp]
        L.lvo_get_field.argtypes = [vp, C.c_char_p, dp]
        L.lvo_set_field.argtypes = [vp, C.c_char_p, dp]
        L.lvo_remesh.argtypes = [vp]
        L.lvo_nnz.restype = C.c_int64
        L.lvo_nnz.argtypes = [vp]
        L.lvo_get_mesh.argtypes = [vp, ip, vp]
        L.lvo_area.argtypes = [vp, dp]
        L.lvo_centroid.argtypes = [vp, dp]
        L.lvo_pressure_assemble.argtypes = [vp, C.c_double]
        L.lvo_pressure_get_operator.argtypes = [vp, ip, ip, dp, dp]
        L.lvo_pressure_matvec.argtypes = [vp, dp, dp]
        L.lvo_pressure_rhs.argtypes = [vp, C.c_double, C.c_int, dp, dp, dp, dp]
        L.lvo_find_pressure.argtypes = [vp, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, dp,
                                        C.POINTER(C.c_int32), dp]
        L.lvo_minres.argtypes = [vp, dp, dp, C.c_double, C.c_double, C.c_int, C.c_int]
        L.lvo_cg.argtypes = [vp, dp, dp, C.c_double, C.c_double, C.c_int]
        L.lvo_populate_hex.argtypes = [vp]
        L.lvo_populate_circ.argtypes = [vp, C.c_double, C.c_double]
        L.lvo_populate_rect.argtypes = [vp]
        L.lvo_populate_vogel.argtypes = [vp, C.c_double, C.c_double]
        L.lvo_populate_rand.argtypes = [vp, dp, C.c_int64]
        L.lvo_lloyd.argtypes = [vp, C.c_int]
        L.lvo_move.argtypes = [vp, C.c_double]
        L.lvo_stiffened_eos.argtypes = [vp, C.c_double, C.c_double]
        L.lvo_ideal_eos.argtypes = [vp, C.c_double, C.c_double]
        L.lvo_pressure_step.argtypes = [vp, C.c_double]
        L.lvo_find_D.argtypes = [vp]
        L.lvo_viscous_step.argtypes = [vp, C.c_double, C.c_int]
        L.lvo_bdary_friction.argtypes = [vp, C.c_double, C.c_void_p]
        L.lvo_boundary_edges.restype = C.c_int64
        L.lvo_boundary_edges.argtypes = [vp, dp, ip, ip]
        L.lvo_set_vbc_edge.argtypes = [vp, dp, C.c_int64]
        L.lvo_bdary_friction_ex.argtypes = [vp, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.lvo_find_dv.argtypes = [vp, C.c_double, C.c_double]
        L.lvo_relaxation_step.argtypes = [vp, C.c_double, C.c_int]
        L.lvo_multiphase_projection.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.lvo_multiphase_apply.argtypes = [vp, dp, dp, dp]
        L.lvo_gravity_step.argtypes = [vp, C.c_double, C.c_double, C.c_double]
        L.lvo_set_threads.argtypes = [C.c_int]
        L.lvo_get_threads.restype = C.c_int
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


_NCOMP = {"x": 2, "rho": 1, "v": 2, "e": 1, "P": 1, "c2": 1, "dv": 2, "mass": 1, "momentum": 2, "energy": 1,
          "phase": 1, "quality": 1, "D": 4, "mu": 1}


def set_threads(n: int) -> None:
    lib().lvo_set_threads(int(n))


def get_threads() -> int:
    return int(lib().lvo_get_threads())


class OracleGrid:
    """VoronoiGrid{PolygonNS} of the reference (voronoigrid.jl:14-50), restated on the CPU."""

    def __init__(self, bmin, bmax, dr, h=None, r_max=None, xperiodic=False, yperiodic=False, full_path=False):
        L = lib()
        self.dr = float(dr)
        self.h = 2.0 * dr if h is None else float(h)          # voronoigrid.jl:27
        self.r_max = 10.0 * dr if r_max is None else float(r_max)
        bmin = np.asarray(bmin, dtype=np.float64)
        bmax = np.asarray(bmax, dtype=np.float64)
        self.bmin, self.bmax = bmin, bmax
        self.xperiodic, self.yperiodic = bool(xperiodic), bool(yperiodic)
        self._g = L.lvo_grid_create(_dp(bmin), _dp(bmax), self.dr, self.h, self.r_max, int(xperiodic),
                                    int(yperiodic), int(full_path))
        if not self._g:
            raise ValueError("h must be positive")  # neighborlist.jl:19-21

    def __del__(self):
        if getattr(self, "_g", None):
            lib().lvo_grid_destroy(self._g)
            self._g = None

    # -- grid / cell list
    def info(self):
        n1, n2, npath = C.c_int64(), C.c_int64(), C.c_int64()
        origin = np.zeros(2)
        lib().lvo_grid_info(self._g, C.byref(n1), C.byref(n2), C.byref(npath), _dp(origin))
        return {"n1": n1.value, "n2": n2.value, "npath": npath.value, "origin": origin}

    def magic_path(self, cap=None):
        cap = self.info()["npath"] if cap is None else cap
        i1 = np.zeros(cap, np.int64)
        i2 = np.zeros(cap, np.int64)
        rr = np.zeros(cap)
        m = lib().lvo_magic_path(self._g, cap, _ip(i1), _ip(i2), _dp(rr))
        return i1[:m], i2[:m], rr[:m]

    def set_rects(self, bmin, bmax, cmin, cmax):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (bmin, bmax, cmin, cmax)]
        lib().lvo_set_rects(self._g, *[_dp(v) for v in a])

    @property
    def n(self):
        return int(lib().lvo_npolygons(self._g))

    def set_points(self, xy):
        xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        lib().lvo_set_points(self._g, xy.shape[0], _dp(xy))

    def get(self, name):
        nc = _NCOMP[name]
        out = np.zeros((self.n, nc) if nc > 1 else self.n)
        st = lib().lvo_get_field(self._g, name.encode(), _dp(out))
        assert st == 0
        return out

    def set(self, name, val):
        nc = _NCOMP[name]
        val = np.ascontiguousarray(np.broadcast_to(np.asarray(val, dtype=np.float64), (self.n, nc) if nc > 1 else (self.n,)))
        st = lib().lvo_set_field(self._g, name.encode(), _dp(val))
        assert st == 0

    # -- mesh
    def remesh(self) -> int:
        return int(lib().lvo_remesh(self._g))

    def mesh(self):
        n = self.n
        nnz = int(lib().lvo_nnz(self._g))
        rowptr = np.zeros(n + 1, np.int64)
        edges = np.zeros(nnz, EDGE_DTYPE)
        lib().lvo_get_mesh(self._g, _ip(rowptr), edges.ctypes.data_as(C.c_void_p))
        return rowptr, edges

    def area(self):
        out = np.zeros(self.n)
        lib().lvo_area(self._g, _dp(out))
        return out

    def centroid(self):
        out = np.zeros((self.n, 2))
        lib().lvo_centroid(self._g, _dp(out))
        return out

    # -- pressure
    def assemble(self, dt):
        lib().lvo_pressure_assemble(self._g, float(dt))

    def operator(self):
        n = self.n
        rowptr = np.zeros(n + 1, np.int64)
        lib().lvo_pressure_get_operator(self._g, _ip(rowptr), None, None, None)
        nnz = int(rowptr[-1])
        col = np.zeros(nnz, np.int64)
        w = np.zeros(nnz)
        diag = np.zeros(n)
        lib().lvo_pressure_get_operator(self._g, _ip(rowptr), _ip(col), _dp(w), _dp(diag))
        return rowptr, col, w, diag

    def matvec(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        lib().lvo_pressure_matvec(self._g, _dp(x), _dp(y))
        return y

    def rhs(self, dt, gp_step=False, vbc_wall=None):
        n = self.n
        b, P0, GP = np.zeros(n), np.zeros(n), np.zeros((n, 2))
        vw = None if vbc_wall is None else np.ascontiguousarray(vbc_wall, dtype=np.float64)
        lib().lvo_pressure_rhs(self._g, float(dt), int(gp_step), _dp(vw), _dp(b), _dp(P0), _dp(GP))
        return b, P0, GP

    def find_pressure(self, dt, niter=10, rtol=1e-6, atol=1e-6, itmax=1000, solver="minres", vbc_wall=None):
        iters = np.zeros(niter, np.int32)
        relres = np.zeros(niter)
        vw = None if vbc_wall is None else np.ascontiguousarray(vbc_wall, dtype=np.float64)
        st = lib().lvo_find_pressure(self._g, float(dt), int(niter), float(rtol), float(atol), int(itmax),
                                     0 if solver == "minres" else 1, _dp(vw),
                                     iters.ctypes.data_as(C.POINTER(C.c_int32)), _dp(relres))
        assert st == 0
        return iters, relres

    def minres(self, b, x0, rtol=1e-6, atol=1e-6, itmax=1000, warm_start=True):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.array(x0, dtype=np.float64, copy=True)
        it = lib().lvo_minres(self._g, _dp(b), _dp(x), rtol, atol, itmax, int(warm_start))
        return x, int(it)

    def cg(self, b, x0, rtol=1e-10, atol=0.0, itmax=100000):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.array(x0, dtype=np.float64, copy=True)
        it = lib().lvo_cg(self._g, _dp(b), _dp(x), rtol, atol, itmax)
        return x, int(it)

    # -- callers either side of the hot path
    def populate_hex(self) -> int:
        return int(lib().lvo_populate_hex(self._g))

    def populate_circ(self, center=(0.0, 0.0)) -> int:
        return int(lib().lvo_populate_circ(self._g, float(center[0]), float(center[1])))

    def populate_rect(self) -> int:
        return int(lib().lvo_populate_rect(self._g))

    def populate_vogel(self, center=(0.0, 0.0)) -> int:
        return int(lib().lvo_populate_vogel(self._g, float(center[0]), float(center[1])))

    def populate_rand(self, samples) -> int:
        """populate_rand! with the caller's uniform samples [[s1, s2], ...] in place of Julia's global RNG."""
        s = np.ascontiguousarray(samples, dtype=np.float64).reshape(-1, 2)
        return int(lib().lvo_populate_rand(self._g, _dp(s), s.shape[0]))

    def lloyd(self, niterations=100) -> int:
        """The relaxation loop of populate_lloyd! (populate.jl:132-145) on the current generators."""
        return int(lib().lvo_lloyd(self._g, int(niterations)))

    def move(self, dt) -> int:
        return int(lib().lvo_move(self._g, float(dt)))

    def stiffened_eos(self, gamma=1.4, P0=0.0):
        lib().lvo_stiffened_eos(self._g, float(gamma), float(P0))

    def ideal_eos(self, gamma=1.4, Pmin=0.0):
        lib().lvo_ideal_eos(self._g, float(gamma), float(Pmin))

    def pressure_step(self, dt):
        lib().lvo_pressure_step(self._g, float(dt))

    def find_D(self):
        lib().lvo_find_D(self._g)

    def viscous_step(self, dt, artificial_viscosity=True):
        lib().lvo_viscous_step(self._g, float(dt), int(artificial_viscosity))

    def bdary_friction(self, dt, vwall=None):
        """diffusion.jl:64-80 with per-wall constant Dirichlet velocities vwall[4][2] (UP, RIGHT, DOWN, LEFT)."""
        vw = np.ascontiguousarray(np.zeros((4, 2)) if vwall is None else vwall, dtype=np.float64)
        lib().lvo_bdary_friction(self._g, float(dt), vw.ctypes.data)

    def boundary_edges(self):
        """(midpoint[nb,2], label[nb], polygon[nb] 1-based) of boundaries(p) over the grid, polygon by polygon."""
        nb = int(lib().lvo_boundary_edges(self._g, None, None, None))
        mid, lab, pol = np.zeros((nb, 2)), np.zeros(nb, np.int64), np.zeros(nb, np.int64)
        if nb:
            lib().lvo_boundary_edges(self._g, _dp(mid), _ip(lab), _ip(pol))
        return mid, lab, pol

    def set_vbc_edge(self, v=None):
        """boundary_velocity(midpoint(e), e.label) per boundary edge (pressure.jl:182), or None for the per-wall constants."""
        if v is None:
            lib().lvo_set_vbc_edge(self._g, None, 0)
        else:
            v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1, 2)
            lib().lvo_set_vbc_edge(self._g, _dp(v), v.shape[0])

    def bdary_friction_ex(self, dt, vwall=None, wall_on=None, v_edge=None, on_edge=None):
        """diffusion.jl:64-80 with vDirichlet(m) / charfun(m) given per boundary edge (or per wall)."""
        def p(a, dt_):
            return None if a is None else np.ascontiguousarray(a, dtype=dt_)
        vw, wo, ve, oe = p(vwall, np.float64), p(wall_on, np.uint8), p(v_edge, np.float64), p(on_edge, np.uint8)
        lib().lvo_bdary_friction_ex(self._g, float(dt), *(None if a is None else a.ctypes.data for a in (vw, wo, ve, oe)))

    def find_dv(self, dt, alpha=1.0):
        lib().lvo_find_dv(self._g, float(dt), float(alpha))

    def relaxation_step(self, dt, rusanov=True) -> int:
        return int(lib().lvo_relaxation_step(self._g, float(dt), int(rusanov)))

    def multiphase_projection(self, quality_threshold=0.25, rtol=1e-4, atol=1e-4, itmax=200):
        it, ok = C.c_int(), C.c_int()
        st = lib().lvo_multiphase_projection(self._g, float(quality_threshold), float(rtol), float(atol), int(itmax), C.byref(it), C.byref(ok))
        assert st == 0
        return it.value, bool(ok.value)

    def multiphase_apply(self, x):
        """(A x, b): the matrix-free projector applied to x (relaxation.jl:91-123) and its right-hand side (:162-177)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y, b = np.zeros_like(x), np.zeros_like(x)
        lib().lvo_multiphase_apply(self._g, _dp(x), _dp(y), _dp(b))
        return y, b

    def gravity_step(self, g, dt):
        lib().lvo_gravity_step(self._g, float(g[0]), float(g[1]), float(dt))
