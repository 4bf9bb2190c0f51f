"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle): the oracle must keep
reproducing them bit for bit (CPU), and the CUDA path must match them through the C ABI (GPU)."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def _load(path):
    d = np.load(path)
    return {k: d[k] for k in d.files}


def test_fixtures_exist():
    assert len(FIXTURES) >= 3


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_oracle_reproduces_golden(oracle, path):
    f = _load(path)
    g = oracle.OracleGrid(tuple(f["bmin"]), tuple(f["bmax"]), float(f["dr"]), xperiodic=bool(f["xper"]), yperiodic=bool(f["yper"]))
    g.set_points(f["xy"])
    assert g.remesh() == 0
    rowptr, edges = g.mesh()
    assert np.array_equal(rowptr, f["rowptr"]) and np.array_equal(edges["label"], f["label"])
    assert edges["v1"].tobytes() == f["v1"].tobytes() and edges["v2"].tobytes() == f["v2"].tobytes()
    assert g.area().tobytes() == f["area"].tobytes() and g.centroid().tobytes() == f["centroid"].tobytes()
    for nm, val in (("rho", f["rho"]), ("mass", f["rho"] * f["area"]), ("c2", 100.0), ("v", f["v"]), ("P", f["P"])):
        g.set(nm, val)
    g.assemble(float(f["dt"]))
    rp, col, w, diag = g.operator()
    assert np.array_equal(col, f["op_col"]) and w.tobytes() == f["op_w"].tobytes() and diag.tobytes() == f["op_diag"].tobytes()
    b, P0, GP = g.rhs(float(f["dt"]), False, f["vbc"])
    assert b.tobytes() == f["b"].tobytes() and GP.tobytes() == f["GP"].tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_gpu_matches_golden(lv, path):
    f = _load(path)
    g = lv.VoronoiGrid(lv.Rectangle(tuple(f["bmin"]), tuple(f["bmax"])), float(f["dr"]), xperiodic=bool(f["xper"]), yperiodic=bool(f["yper"]))
    g.set_points(f["xy"])
    lv.remesh(g)
    assert np.array_equal(g.rowptr, f["rowptr"]) and np.array_equal(g.edges["label"], f["label"])       # connectivity bit-exact
    assert g.edges["v1"].tobytes() == f["v1"].tobytes() and g.edges["v2"].tobytes() == f["v2"].tobytes()
    assert np.allclose(lv.area(g), f["area"], rtol=1e-12, atol=0) and np.allclose(lv.centroid(g), f["centroid"], rtol=1e-12, atol=1e-15)
    g.rho[...] = f["rho"]; g.mass[...] = f["rho"] * f["area"]; g.c2[...] = 100.0; g.v[...] = f["v"]; g.P[...] = f["P"]
    s = lv.PressureSolver(g)
    s.upload_fields(g.mass, g.rho, g.c2, g.P, g.v)
    s.assemble(float(f["dt"]))
    rp, col, w, diag = s.operator()
    assert np.array_equal(col, f["op_col"]) and np.allclose(w, f["op_w"], rtol=1e-13, atol=0) and np.allclose(diag, f["op_diag"], rtol=1e-13)
    b, GP = s.rhs(float(f["dt"]), False, f["vbc"])
    assert np.allclose(b, f["b"], rtol=1e-12, atol=1e-12 * np.abs(f["b"]).max())
    x, iters, relres = s.solve(b, f["P"], rtol=1e-13, atol=0.0, itmax=100000)
    assert relres <= 1e-10
    assert np.abs(x - f["P_solved"]).max() <= 1e-8 * np.abs(f["P_solved"]).max()                        # north_star pressure tolerance
