"""Host-side widening rows: VTK / pvd / csv writers on CPU; seeding strategies, Lloyd on
the device and the reference's run! loop (with the file checks of tests/taylorgreen.jl:115-119) on the GPU."""
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest


class _FakeGrid:
    pass


def test_vtp_writers_roundtrip(lv, tmp_path):
    g = _FakeGrid()
    # two triangles
    g.rowptr = np.array([0, 3, 6])
    g.edges = np.zeros(6, lv.EDGE_DTYPE)
    g.edges["v1"] = [[0, 0], [0, 1], [1, 0], [1, 1], [1, 0], [0, 1]]
    g.x = np.array([[0.3, 0.3], [0.7, 0.7]])
    g.P = np.array([1.5, -2.0]); g.v = np.array([[1.0, 2.0], [3.0, 4.0]])
    f = lv.io.export_grid(g, str(tmp_path / "c"), "P", "v")
    root = ET.parse(f).getroot()
    piece = root.find("PolyData/Piece")
    assert piece.get("NumberOfPoints") == "6" and piece.get("NumberOfPolys") == "2"
    offs = [int(t) for t in piece.find("Polys/DataArray[@Name='offsets']").text.split()]
    assert offs == [3, 6]
    P = [float(t) for t in piece.find("CellData/DataArray[@Name='P']").text.split()]
    assert P == [1.5, -2.0]
    v = piece.find("CellData/DataArray[@Name='v']")
    assert v.get("NumberOfComponents") == "3"
    f2 = lv.io.export_points(g, str(tmp_path / "p"), "P")
    assert ET.parse(f2).getroot().find("PolyData/Piece").get("NumberOfVerts") == "2"
    with pytest.raises(ValueError, match="does not exist"):                      # IO.jl:65-67
        lv.io.export_grid(g, str(tmp_path / "bad"), "nope")


@pytest.mark.gpu
def test_populate_strategies_and_lloyd(lv):
    dom = lv.Rectangle((0.0, 0.0), (1.0, 1.0))
    dr = 1 / 40
    counts = {}
    for name in ("rect", "hex", "rand", "circ", "vogel"):
        g = lv.VoronoiGrid(dom, dr)
        kw = {"center": (0.5, 0.5)} if name in ("circ", "vogel") else {}
        getattr(lv.populate, "populate_" + name)(g, **kw)
        counts[name] = g.n
        assert abs(lv.area(g).sum() - 1.0) < 1e-12
        assert abs(g.n - 1600) < 200
    g = lv.VoronoiGrid(dom, dr)
    lv.populate.populate_hex(g)
    inner = (np.abs(g.x - 0.5) < 0.35).all(1)
    assert np.allclose(lv.area(g)[inner], dr * dr, rtol=1e-12)                   # a*b = dr^2  populate.jl:156-157
    # Lloyd relaxation evens the cells out: area spread shrinks a lot compared with the random seeding
    g1 = lv.VoronoiGrid(dom, dr); lv.populate.populate_rand(g1, seed=3)
    g2 = lv.VoronoiGrid(dom, dr); lv.populate.populate_lloyd(g2, niterations=30, seed=3)
    assert g1.n == g2.n
    assert lv.area(g2).std() < 0.35 * lv.area(g1).std()
    assert abs(lv.area(g2).sum() - 1.0) < 1e-12


@pytest.mark.gpu
def test_run_loop_writes_the_reference_files(lv, tmp_path):
    """tests/taylorgreen.jl:101-119: run! with nframes = 20, save_points, vtp_vars = (P, v) must leave simdata.csv,
    cells.pvd, cframe19.vtp, points.pvd, pframe19.vtp."""
    S = lv.stepping
    N = 32; Re = 400.0; dr = 1.0 / N; dt = 0.1 * dr; gamma = 1.4; P0 = 50.0 ** 2 / gamma

    class Sim:
        pass

    sim = Sim()
    sim.grid = g = lv.VoronoiGrid(lv.Rectangle((0, 0), (1, 1)), dr, xperiodic=True, yperiodic=True)

    def ic(grid):
        v, P = lv.synthetic.taylor_green_fields(grid.x, 0.0, Re)
        grid.v[...] = v; grid.rho[...] = 1.0; grid.mass[...] = lv.area(grid); grid.P[...] = P
        grid.e[...] = 0.5 * (v ** 2).sum(1) + P / (gamma - 1.0); grid.mu[...] = 1.0 / Re

    lv.populate.populate_hex(g, ic=ic)
    S.to_device(g)
    sim.solver = lv.PressureSolver(g)
    sim.E = 0.0

    def step(sim, t):
        S.move(g, dt); S.stiffened_eos(g, gamma, P0); S.find_pressure_resident(sim.solver, dt)
        S.pressure_step(g, dt); S.find_D(g); S.viscous_step(g, dt, False); S.find_dv(g, dt); S.relaxation_step(g, dt)

    def postproc(sim, t):
        sim.E = float((g.mass * g.e).sum())

    out = str(tmp_path / "results")
    lv.io.run(sim, dt, 20 * dt, step, nframes=20, path=out, save_csv=True, save_points=True, save_grid=True, vtp_vars=("P", "v"),
              csv_vars=("E",), postproc=postproc, sync=lambda s: S.from_device(g, mesh=True))
    for f in ("simdata.csv", "cells.pvd", "cframe19.vtp", "points.pvd", "pframe19.vtp"):
        assert os.path.exists(os.path.join(out, f)), f
    rows = open(os.path.join(out, "simdata.csv")).read().strip().splitlines()
    assert rows[0] == "time,E" and len(rows) == 21
    E = np.array([float(r.split(",")[1]) for r in rows[1:]])
    assert np.abs(E - E[0]).max() < 1e-10                                        # energy is conserved to rounding
