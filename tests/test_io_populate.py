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
    for mode, kw in (("ascii", {"ascii": True}), ("raw", {}), ("zlib", {"compress": True})):
        f = lv.io.export_grid(g, str(tmp_path / ("c" + mode)), "P", "v", **kw)
        d = lv.io.read_vtp(f)
        assert d["_counts"]["Points"] == 6 and d["_counts"]["Polys"] == 2
        assert d["offsets"].tolist() == [3, 6] and d["connectivity"].tolist() == list(range(6))
        assert np.array_equal(d["Points"], np.column_stack([g.edges["v1"], np.zeros(6)]))
        assert d["P"].tolist() == [1.5, -2.0]
        assert np.array_equal(d["v"], [[1.0, 2.0, 0.0], [3.0, 4.0, 0.0]])           # Vec3  IO.jl:1-3
        f2 = lv.io.export_points(g, str(tmp_path / ("p" + mode)), "P", **kw)
        assert lv.io.read_vtp(f2)["_counts"]["Verts"] == 2
        if mode == "ascii":                                                          # the ascii form is plain XML
            piece = ET.parse(f).getroot().find("PolyData/Piece")
            assert piece.get("NumberOfPoints") == "6" and piece.find("CellData/DataArray[@Name='v']").get("NumberOfComponents") == "3"
        else:                                                                        # the XML part parses once the raw block is cut out
            blob = open(f, "rb").read()
            head = blob[:blob.index(b"<AppendedData")] + b"</VTKFile>"
            assert ET.fromstring(head).find("PolyData/Piece/Points/DataArray").get("format") == "appended"
    # a large frame goes through without per-number formatting: 2M edge records in well under a second of CPU
    import time
    big = _FakeGrid()
    m = 2_000_000
    big.rowptr = np.arange(0, m + 1, 4)
    big.edges = np.zeros(m, lv.EDGE_DTYPE)
    big.edges["v1"] = np.random.default_rng(0).random((m, 2))
    big.P = np.zeros(m // 4)
    t0 = time.perf_counter()
    fb = lv.io.export_grid(big, str(tmp_path / "big"), "P")
    assert time.perf_counter() - t0 < 5.0
    assert np.array_equal(lv.io.read_vtp(fb)["Points"][:, :2], big.edges["v1"])
    with pytest.raises(ValueError, match="does not exist"):                      # IO.jl:65-67
        lv.io.export_grid(g, str(tmp_path / "bad"), "nope")


def _host_points(lv, name, dom, dr, **kw):
    """The generator set a populate_* strategy of the host mirror produces, without touching the GPU."""
    class Dummy:                                                            # stands in for VoronoiGrid: only geometry is read
        boundary_rect, n = dom, 0
    d = Dummy(); d.dr = dr
    got = {}
    orig = lv.populate._finish
    lv.populate._finish = lambda grid, pts, charfun, ic, edges=True: got.setdefault("pts", _inside(np.asarray(pts, float).reshape(-1, 2), dom))
    try:
        getattr(lv.populate, "populate_" + name)(d, **kw)
    finally:
        lv.populate._finish = orig
    return got["pts"]


def _inside(p, dom):
    (x0, y0), (x1, y1) = dom.xmin, dom.xmax
    return p[(p[:, 0] >= x0) & (p[:, 0] <= x1) & (p[:, 1] >= y0) & (p[:, 1] <= y1)]


@pytest.mark.parametrize("dom,dr", [(((0.0, 0.0), (1.0, 1.0)), 1 / 40), (((-0.5, 0.25), (1.5, 1.25)), 1 / 37), (((-1.0, -1.0), (1.0, 1.0)), 0.0625)])
def test_seeding_strategies_match_the_oracle_restatement(lv, oracle, dom, dr):
    """populate_rect! / hex! / circ! / vogel! / rand! (populate.jl:18-174): the host mirror and the C restatement generate the
    same generators in the same order (labels are positions in that order).  rect and hex are exact; circ and vogel go
    through libm's sin / cos on both sides and agree to the last few ulps."""
    rect = lv.Rectangle(*dom)
    centre = (0.5 * (dom[0][0] + dom[1][0]) + 0.013, 0.5 * (dom[0][1] + dom[1][1]) - 0.02)
    samples = np.random.default_rng(5).random((int(round((dom[1][0] - dom[0][0]) * (dom[1][1] - dom[0][1]) / dr ** 2)), 2))
    cases = {"rect": ({}, lambda og: og.populate_rect(), 0.0), "hex": ({}, lambda og: og.populate_hex(), 0.0),
             "circ": ({"center": centre}, lambda og: og.populate_circ(centre), 4e-16),
             "vogel": ({"center": centre}, lambda og: og.populate_vogel(centre), 4e-16),
             "rand": ({"samples": samples}, lambda og: og.populate_rand(samples), 0.0)}
    for name, (kw, run, tol) in cases.items():
        og = oracle.OracleGrid(dom[0], dom[1], dr)
        assert run(og) == 0, name
        a, b = _host_points(lv, name, rect, dr, **kw), og.get("x")
        assert a.shape == b.shape and len(a) > 100, (name, a.shape, b.shape)
        scale = np.abs(b).max()
        assert np.abs(a - b).max() <= tol * scale, (name, np.abs(a - b).max())


def test_lloyd_relaxation_on_the_restatement(oracle):
    """populate_lloyd! (populate.jl:132-145) through the restatement: the cells even out (area spread shrinks), the domain
    stays tiled, generators converge towards their centroids."""
    dr = 1 / 24
    og = oracle.OracleGrid((0.0, 0.0), (1.0, 1.0), dr)
    s = np.random.default_rng(3).random((576, 2))
    assert og.populate_rand(s) == 0
    a0 = og.area().std()
    assert og.lloyd(30) == 0
    assert og.area().std() < 0.35 * a0 and abs(og.area().sum() - 1.0) < 1e-12
    assert np.abs(og.get("x") - og.centroid()).max() < 0.05 * dr


@pytest.mark.gpu
def test_seeding_and_lloyd_parity_with_the_oracle(lv, oracle):
    """f3 parity: every seeding strategy gives the oracle's mesh byte for byte on the oracle's generators, and the device-side
    Lloyd loop (lv_step_lloyd: 100 x (remesh!; x = centroid)) lands on the restatement's positions bit for bit -- every
    iteration clips the same polygons and computes the same centroids."""
    dom = ((0.0, 0.0), (1.0, 1.0))
    dr = 1 / 32
    centre = (0.47, 0.52)
    samples = np.random.default_rng(8).random((1024, 2))
    runs = {"rect": lambda og: og.populate_rect(), "hex": lambda og: og.populate_hex(), "circ": lambda og: og.populate_circ(centre),
            "vogel": lambda og: og.populate_vogel(centre), "rand": lambda og: og.populate_rand(samples)}
    for name, run in runs.items():
        og = oracle.OracleGrid(dom[0], dom[1], dr)
        assert run(og) == 0
        g = lv.VoronoiGrid(lv.Rectangle(*dom), dr)
        g.set_points(og.get("x"))
        lv.remesh(g)
        r0, e0 = og.mesh()
        assert np.array_equal(g.rowptr, r0) and g.edges.tobytes() == e0.tobytes(), name
        assert np.array_equal(lv.centroid(g), og.centroid()) and np.array_equal(lv.area(g), og.area()), name
    # Lloyd from the same random seeding, 100 iterations like the reference's default
    og = oracle.OracleGrid(dom[0], dom[1], dr)
    assert og.populate_rand(samples) == 0
    g = lv.VoronoiGrid(lv.Rectangle(*dom), dr)
    lv.populate.populate_lloyd(g, niterations=100, samples=samples)
    assert og.lloyd(100) == 0
    assert np.array_equal(g.x, og.get("x"))
    r0, e0 = og.mesh()
    assert np.array_equal(g.rowptr, r0) and g.edges.tobytes() == e0.tobytes()


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
