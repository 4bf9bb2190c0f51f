"""GPU parity tests of the device-resident explicit sweeps against the oracle, and the
reference's own integration test (tests/taylorgreen.jl) run end to end on the GPU path."""
import os

import numpy as np
import pytest

from .conftest import make_points

pytestmark = pytest.mark.gpu

FIELDS = ["x", "v", "rho", "e", "P", "c2", "mass", "dv", "quality", "momentum", "energy"]


def _pair(lv, oracle, kind, n_side, xper, yper, seed):
    xy, dr, bmin, bmax = make_points(kind, n_side, seed)
    rng = np.random.default_rng(seed)
    n = len(xy)
    og = oracle.OracleGrid(bmin, bmax, dr, xperiodic=xper, yperiodic=yper)
    og.set_points(xy); assert og.remesh() == 0
    g = lv.VoronoiGrid(lv.Rectangle(bmin, bmax), dr, xperiodic=xper, yperiodic=yper)
    g.set_points(xy)
    v, P = lv.synthetic.taylor_green_fields(xy)
    v = v + 0.05 * rng.standard_normal((n, 2))
    rho = 1.0 + 0.3 * rng.random(n)
    mass = rho * og.area()
    e = 0.5 * (v ** 2).sum(1) + (P + 50.0) / (rho * 0.4)
    mu = np.full(n, 1.0 / 400)
    for name, val in (("v", v), ("rho", rho), ("mass", mass), ("e", e), ("P", P), ("mu", mu), ("c2", 100.0 + 0 * rho)):
        og.set(name, val)
        getattr(g, name)[...] = val
    lv.stepping.to_device(g)
    return g, og, dr


def _compare(lv, g, og, names, rtol=1e-12):
    for nm in names:
        a = lv.stepping.state_get(g, nm)
        b = og.get(nm)
        scale = np.abs(b).max() + 1e-300
        assert np.abs(a - b).max() <= rtol * scale, (nm, np.abs(a - b).max() / scale)


@pytest.mark.parametrize("kind,n_side,xper,yper", [("jitter", 48, True, True), ("poisson", 40, False, False), ("rect2x1", 32, True, False)])
def test_explicit_sweeps_match_oracle(lv, oracle, kind, n_side, xper, yper):
    S = lv.stepping
    g, og, dr = _pair(lv, oracle, kind, n_side, xper, yper, 9)
    dt = 0.1 * dr
    S.stiffened_eos(g, 1.4, 30.0); og.stiffened_eos(1.4, 30.0)
    _compare(lv, g, og, ["rho", "P", "c2"])
    S.pressure_step(g, dt); og.pressure_step(dt)
    _compare(lv, g, og, ["v", "e"])
    S.find_D(g); og.find_D()
    assert np.allclose(S.state_get(g, "D"), og.get("D"), rtol=1e-12, atol=1e-12 * np.abs(og.get("D")).max())
    S.viscous_step(g, dt, True); og.viscous_step(dt, True)
    _compare(lv, g, og, ["v", "e"])
    vwall = np.array([[1.0, 0.0], [0.0, -0.5], [0.25, 0.0], [0.0, 0.0]])     # lid + moving side walls (cavity.jl:41-44)
    S.bdary_friction(g, dt, vwall); og.bdary_friction(dt, vwall)
    _compare(lv, g, og, ["v", "e"])
    S.find_dv(g, dt, 1.0); og.find_dv(dt, 1.0)
    _compare(lv, g, og, ["dv", "quality"])
    S.relaxation_step(g, dt, True); assert og.relaxation_step(dt, True) == 0
    _compare(lv, g, og, ["x", "mass", "v", "e", "momentum", "energy"])
    S.ideal_eos(g, 1.4, 0.1); og.ideal_eos(1.4, 0.1)
    _compare(lv, g, og, ["rho", "P", "c2"])
    S.move(g, dt); assert og.move(dt) == 0
    _compare(lv, g, og, ["x", "v"])
    # the mesh after the device-side moves is still the oracle's, bit for bit
    rowptr, edges, area, cen = g.mesh_download(g.n)
    r0, e0 = og.mesh()
    assert np.array_equal(rowptr, r0) and edges.tobytes() == e0.tobytes()


def test_bdary_friction_closures(lv, oracle):
    """bdary_friction!(grid, vDirichlet, dt; charfun) (diffusion.jl:64-80) with the reference's closures: a lid velocity that
    varies along the wall and charfun = top_and_bottom (examples/bubble.jl) -- side walls get NO drag, not zero-velocity drag."""
    S = lv.stepping
    g, og, dr = _pair(lv, oracle, "poisson", 40, False, False, 9)
    dt = 0.3 * dr
    mu = np.full(g.n, 2.0e-2)
    og.set("mu", mu); S.state_set(g, "mu", mu)
    vD = lambda m: np.array([np.sin(np.pi * m[0]) ** 2, 0.1 * m[1]])       # noqa: E731
    top_and_bottom = lambda m: (m[1] > 1.0 - 1e-9) or (m[1] < 1e-9)          # noqa: E731
    mid, lab, _ = og.boundary_edges()
    v0 = og.get("v").copy()
    S.bdary_friction(g, dt, vD, charfun=top_and_bottom)
    og.bdary_friction_ex(dt, v_edge=np.array([vD(m) for m in mid]), on_edge=np.array([top_and_bottom(m) for m in mid], dtype=np.uint8))
    _compare(lv, g, og, ["v", "e"])
    v1 = og.get("v")
    x = og.get("x")
    side_only = ((x[:, 0] < 0.5 * dr) | (x[:, 0] > 1 - 0.5 * dr)) & (x[:, 1] > 3 * dr) & (x[:, 1] < 1 - 3 * dr)
    assert side_only.any() and np.array_equal(v1[side_only], v0[side_only])   # cells touching only a side wall are untouched
    assert np.abs(v1 - v0).max() > 0
    # per-wall flags: the same charfun as four switches (UP, RIGHT, DOWN, LEFT)
    S.bdary_friction(g, dt, np.array([[1.0, 0.0], [0, 0], [0, 0], [0, 0]]), charfun=[1, 0, 1, 0])
    og.bdary_friction_ex(dt, vwall=np.array([[1.0, 0.0], [0, 0], [0, 0], [0, 0]]), wall_on=np.array([1, 0, 1, 0], dtype=np.uint8))
    _compare(lv, g, og, ["v", "e"])


def test_walls_stop_escaping_generators(lv, oracle):
    """move! projects the velocity of cells that would leave the box and zeroes it if that fails (move.jl:9-21)."""
    S = lv.stepping
    g, og, dr = _pair(lv, oracle, "poisson", 32, False, False, 4)
    x0 = og.get("x")
    vv = 0.2 * og.get("v")
    near = (x0[:, 0] > 1.0 - 1.5 * dr) | (x0[:, 1] < 1.5 * dr)      # cells at the right / bottom wall run into it
    vv[near] = np.array([3.0, -2.0])
    og.set("v", vv); S.state_set(g, "v", vv)
    dt = 0.6 * dr
    S.move(g, dt); assert og.move(dt) == 0
    _compare(lv, g, og, ["x", "v"])
    x = S.state_get(g, "x")
    assert x.min() >= 0.0 and x.max() <= 1.0
    v1 = S.state_get(g, "v")
    assert (np.abs(v1[near] - np.array([3.0, -2.0])).max(1) > 0).any()   # some velocities were projected or zeroed


def test_taylor_green_on_the_gpu(lv, oracle):
    """The reference's only test (tests/taylorgreen.jl): N = 80, Re = 400, 80 steps of the canonical step!, thresholds
    E_err < 1e-8, v_err < 0.01, P_err < 0.01 (:112-114) -- every operator of the step on the GPU."""
    S = lv.stepping
    N = 80; Re = 400.0; rho0 = 1.0; dr = 1.0 / N; dt = 0.1 * dr; c0 = 50.0; gamma = 1.4; t_end = 0.1
    P0 = rho0 * c0 ** 2 / gamma
    og = oracle.OracleGrid((0, 0), (1, 1), dr, xperiodic=True, yperiodic=True)
    assert og.populate_hex() == 0                                   # seeding only (populate.jl:149-174)
    xy = og.get("x")
    g = lv.VoronoiGrid(lv.Rectangle((0, 0), (1, 1)), dr, xperiodic=True, yperiodic=True)
    g.set_points(xy)
    lv.remesh(g, edges=False)
    area = lv.area(g).copy()
    v, P = lv.synthetic.taylor_green_fields(xy, 0.0, Re)
    g.v[...] = v; g.rho[...] = rho0; g.mass[...] = rho0 * area; g.P[...] = P
    g.e[...] = 0.5 * (v ** 2).sum(1) + P / (rho0 * (gamma - 1.0)); g.mu[...] = 1.0 / Re
    S.to_device(g)
    solver = lv.PressureSolver(g)
    t = 0.0; k = 0; E0 = None; errs = None
    k_frame = max(round(t_end / (20 * dt)), 1)
    while t < t_end:
        k += 1
        S.move(g, dt)
        S.stiffened_eos(g, gamma, P0)
        S.find_pressure_resident(solver, dt)
        S.pressure_step(g, dt); S.find_D(g); S.viscous_step(g, dt, False); S.find_dv(g, dt)
        S.relaxation_step(g, dt)
        if k % k_frame == 0:
            S.from_device(g)
            _, _, area, _ = g.mesh_download(g.n, edges=False)
            p_avg = (area * g.P).sum()
            E = (g.mass * g.e).sum()
            E0 = E if E0 is None else E0
            ve, Pe = lv.synthetic.taylor_green_fields(g.x, t, Re)
            errs = (E - E0, np.sqrt((area * ((g.v - ve) ** 2).sum(1)).sum()), np.sqrt((area * (g.P - p_avg - Pe) ** 2).sum()))
        t += dt
    assert k == 80
    E_err, v_err, P_err = errs
    assert E_err < 1e-8 and v_err < 0.01 and P_err < 0.01, errs


def _two_phase(lv, oracle, n_side=40):
    """A Rayleigh-Taylor style set-up (examples/rayleightaylor.jl:55-62): heavy fluid above a wavy interface, walls in y."""
    xy, dr, bmin, bmax = make_points("jitter", n_side, 2)
    n = len(xy)
    og = oracle.OracleGrid(bmin, bmax, dr, xperiodic=True, yperiodic=False)
    og.set_points(xy); assert og.remesh() == 0
    g = lv.VoronoiGrid(lv.Rectangle(bmin, bmax), dr, xperiodic=True, yperiodic=False)
    g.set_points(xy)
    up = xy[:, 1] > 0.5 + 0.05 * np.cos(2 * np.pi * xy[:, 0])
    phase = np.where(up, 0.0, 1.0)
    rho = np.where(up, 1.8, 1.0)
    area = og.area()
    v = np.zeros((n, 2)); P = 10.0 - rho * 0.5 * xy[:, 1]
    e = 0.5 * (v ** 2).sum(1) + P / (rho * 0.4)
    for name, val in (("phase", phase), ("rho", rho), ("mass", rho * area), ("v", v), ("P", P), ("e", e), ("mu", 1e-3 + 0 * rho)):
        og.set(name, val)
        getattr(g, name)[...] = val
    lv.stepping.to_device(g)
    return g, og, dr


def test_multiphase_projection_matches_oracle(lv, oracle):
    """multiphase_projection! (relaxation.jl:179-206): same MINRES on the same matrix-free projector on both sides."""
    S = lv.stepping
    oracle.set_threads(1)                                               # OpenMP reductions are order-nondeterministic; MINRES on a
    g, og, dr = _two_phase(lv, oracle)                                  # singular operator is sensitive to that
    rng = np.random.default_rng(1)
    dv = 0.05 * rng.standard_normal((g.n, 2))
    q = 0.1 + 0.9 * rng.random(g.n)                                   # some cells below the 0.25 quality threshold
    for nm, val in (("dv", dv), ("quality", q)):
        og.set(nm, val); S.state_set(g, nm, val)
    it, ok = S.multiphase_projection(g)                                 # reference settings: 1e-4, 1e-4, 200
    it0, ok0 = og.multiphase_projection()
    assert abs(it - it0) <= 4 and ok == ok0
    a, b = S.state_get(g, "dv"), og.get("dv")
    assert np.abs(a - b).max() <= 1e-3 * np.abs(b).max()               # both stop at the loose reference tolerance (1e-4)
    assert np.abs(b - dv).max() > 1e-3                                  # the projection did change dv at the interface
    low = q < 0.25
    assert np.array_equal(a[low], dv[low])                              # poor cells are left alone (relaxation.jl:194-196)
    # tight tolerances: converged projections agree closely
    for nm, val in (("dv", dv),):
        og.set(nm, val); S.state_set(g, nm, val)
    it1, _ = S.multiphase_projection(g, rtol=1e-11, atol=1e-11, itmax=5000)
    it2, _ = og.multiphase_projection(rtol=1e-11, atol=1e-11, itmax=5000)
    a, b = S.state_get(g, "dv"), og.get("dv")
    err = np.abs(a - b).max() / np.abs(b).max()
    print("multiphase tight:", it1, it2, err)
    # the projector is singular (rows of cells away from the interface vanish); Krylov's MINRES then stops on its
    # conditioning / forward-error estimates, both sides alike, a few digits short of the requested tolerance
    oracle.set_threads(0)
    assert abs(it1 - it2) <= 8 and err <= 1e-5


def test_rayleigh_taylor_steps_match_oracle(lv, oracle):
    """The step! of examples/rayleightaylor.jl:90-102 (gravity, multiphase projector) for a few steps, GPU vs oracle."""
    S = lv.stepping
    g, og, dr = _two_phase(lv, oracle, 32)
    dt = 0.05 * dr
    solver = lv.PressureSolver(g, rtol=1e-12, atol=0.0, itmax=20000)
    for _ in range(3):
        S.move(g, dt); assert og.move(dt) == 0
        S.gravity_step(g, (0.0, -1.0), dt); og.gravity_step((0.0, -1.0), dt)
        S.ideal_eos(g, 1.4, 0.0); og.ideal_eos(1.4, 0.0)
        S.find_pressure_resident(solver, dt); og.find_pressure(dt, 10, rtol=1e-12, atol=0.0, itmax=20000, solver="cg")
        S.pressure_step(g, dt); og.pressure_step(dt)
        S.find_D(g); og.find_D()
        S.viscous_step(g, dt); og.viscous_step(dt)
        S.find_dv(g, dt); og.find_dv(dt)
        S.multiphase_projection(g, rtol=1e-11, atol=1e-11, itmax=5000); og.multiphase_projection(rtol=1e-11, atol=1e-11, itmax=5000)
        S.relaxation_step(g, dt); assert og.relaxation_step(dt) == 0
    for nm in ("x", "v", "e", "mass"):
        a, b = S.state_get(g, nm), og.get(nm)
        assert np.abs(a - b).max() <= 1e-7 * np.abs(b).max(), nm


def test_multiphase_projector_operator_matches_oracle(lv, oracle):
    """mul!(res, A::MultiphaseProjector, x) (relaxation.jl:91-123) and the right-hand side of refresh! (:162-177), applied
    directly -- no Krylov iteration in between -- GPU vs oracle at rounding level; the operator is symmetric."""
    S = lv.stepping
    g, og, dr = _two_phase(lv, oracle)
    rng = np.random.default_rng(5)
    dv = 0.05 * rng.standard_normal((g.n, 2))
    og.set("dv", dv); S.state_set(g, "dv", dv)
    x = rng.standard_normal(g.n)
    y, b = S.multiphase_apply(g, x)
    y0, b0 = og.multiphase_apply(x)
    assert np.abs(y0).max() > 0 and np.abs(b0).max() > 0
    assert np.abs(y - y0).max() <= 1e-12 * np.abs(y0).max()
    assert np.abs(b - b0).max() <= 1e-12 * np.abs(b0).max()
    z = rng.standard_normal(g.n)
    yz, _ = S.multiphase_apply(g, z)
    assert abs(np.dot(z, y) - np.dot(x, yz)) <= 1e-10 * (np.abs(y).max() * np.abs(z).sum())   # <z, A x> = <x, A z>


def test_gresho_vortex_on_the_gpu(lv, oracle):
    """examples/gresho.jl (BASELINE config 0: walls, populate_circ!, ideal EOS at Ma = 0.1, artificial viscosity) on the
    GPU path: 30 steps tracked against the oracle operator for operator, then on to t = 0.25 where the steady exact
    solution bounds the mass-weighted L2 velocity error (gresho.jl:121-126) and energy is conserved."""
    from . import gresho_case as G
    S = lv.stepping
    N = 50
    dr = 1.0 / N
    dt = 0.1 * dr
    g = lv.VoronoiGrid(lv.Rectangle(G.BMIN, G.BMAX), dr)
    lv.populate.populate_circ(g)                                          # populate_circ!  populate.jl:18-35
    ref_pts = G.circ_points(dr)                                           # the oracle-side case accumulates r += dr: ulp-level differences
    assert g.x.shape == ref_pts.shape and np.abs(g.x - ref_pts).max() < 1e-14
    og = oracle.OracleGrid(G.BMIN, G.BMAX, dr)
    og.set_points(g.x); assert og.remesh() == 0
    for k, val in G.initial_fields(g.x, lv.area(g).copy()).items():
        getattr(g, k)[...] = val
        og.set(k, val)
    E0 = (g.mass * g.e).sum()
    S.to_device(g)
    solver = lv.PressureSolver(g, rtol=1e-12, atol=0.0, itmax=20000)

    def step_gpu(sol):                                                    # step!  gresho.jl:101-114
        S.move(g, dt); S.ideal_eos(g, G.GAMMA, 0.0); S.find_pressure_resident(sol, dt)
        S.pressure_step(g, dt); S.find_D(g); S.viscous_step(g, dt, True); S.find_dv(g, dt, 1.0); S.relaxation_step(g, dt, True)

    for _ in range(30):
        step_gpu(solver)
        assert og.move(dt) == 0
        og.ideal_eos(G.GAMMA, 0.0); og.find_pressure(dt, 10, rtol=1e-12, atol=0.0, itmax=20000, solver="cg")
        og.pressure_step(dt); og.find_D(); og.viscous_step(dt, True); og.find_dv(dt, 1.0)
        assert og.relaxation_step(dt, True) == 0
    for nm in ("x", "v", "e", "rho"):
        a, b = S.state_get(g, nm), og.get(nm)
        assert np.abs(a - b).max() <= 1e-7 * np.abs(b).max(), nm
    loose = lv.PressureSolver(g)                                          # reference tolerances for the rest of the run
    for _ in range(round(0.25 / dt) - 30):
        step_gpu(loose)
    S.from_device(g)
    assert abs((g.mass * g.e).sum() - E0) < 1e-11 * abs(E0)
    assert G.l2_error(g.x, g.v, g.mass) < 0.05


def test_sedov_blast_on_the_gpu(lv, oracle):
    """examples/sedov.jl (BASELINE config) at N = 40, every operator of the step on the GPU, pressure solve by the
    reference's Krylov method (MINRES, reference tolerances): same comparison with the reference's semi-analytic profile as
    tests/test_oracle.py runs on the restatement.  (At atol = rtol = 1e-6 the stopping point -- hence the noisy wake --
    depends on the Krylov method: CG gives a wake error of 13 % on the GPU and on the CPU restatement alike, MINRES 8 %.)"""
    from . import sedov_case as C
    S = lv.stepping
    N = 40
    dr = 1.0 / N
    og = oracle.OracleGrid((-1.0, -1.0), (1.0, 1.0), dr)
    assert og.populate_hex() == 0                                       # seeding only (populate.jl:149-174)
    xy = og.get("x")
    g = lv.VoronoiGrid(lv.Rectangle((-1.0, -1.0), (1.0, 1.0)), dr)
    g.set_points(xy)
    lv.remesh(g, edges=False)
    for k, val in C.initial_fields(xy, lv.area(g).copy()).items():
        getattr(g, k)[...] = val
    E0 = (g.mass * g.e).sum()
    S.to_device(g)
    solver = lv.PressureSolver(g, solver="minres")
    for dt in C.time_steps(dr):
        S.move(g, dt)
        S.ideal_eos(g, C.GAMMA, C.P0)
        S.find_pressure_resident(solver, dt)
        S.pressure_step(g, dt); S.find_D(g); S.viscous_step(g, dt, True); S.find_dv(g, dt)
        S.relaxation_step(g, dt)
    S.from_device(g)
    assert abs((g.mass * g.e).sum() - E0) < 1e-13
    c = C.compare_with_reference(g.x, g.rho, dr)
    assert abs(c["r_shock"] - c["r_shock_ref"]) <= 2.0 * dr, c
    assert c["peak"] > 3.0 and c["wake_rel_err"] < 0.10 and c["ahead_err"] < 1e-3, c
