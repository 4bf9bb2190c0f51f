"""Regenerates the golden fixtures in this directory from the CPU oracle (oracle/lv_oracle.c).

The reference ships no golden vectors for the mesh-and-pressure path and cannot run here (no Julia), so these
fixtures are outputs of the restatement, frozen at the commit that first passed the reference's own Taylor-Green
thresholds (tests/test_oracle.py).  They pin the oracle against regressions and let the GPU tests check the CUDA
path against committed numbers.  Usage:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
import lvb200 as lv  # noqa: E402
from tests.conftest import make_points  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "jitter16_periodic": ("jitter", 16, True, True, 0),
    "poisson12_walls": ("poisson", 12, False, False, 1),
    "rect2x1_8_xperiodic": ("rect2x1", 8, True, False, 2),
}


def make(name):
    kind, n_side, xper, yper, seed = CASES[name]
    xy, dr, bmin, bmax = make_points(kind, n_side, seed)
    g = orc.OracleGrid(bmin, bmax, dr, xperiodic=xper, yperiodic=yper)
    g.set_points(xy)
    assert g.remesh() == 0
    rowptr, edges = g.mesh()
    area, cen = g.area(), g.centroid()
    v, P = lv.synthetic.taylor_green_fields(xy)
    rho = np.where(xy[:, 0] > 0.5 * (bmin[0] + bmax[0]), 2.0, 1.0)
    dt = 0.1 * dr
    for nm, val in (("rho", rho), ("mass", rho * area), ("c2", 100.0), ("v", v), ("P", P)):
        g.set(nm, val)
    g.assemble(dt)
    op_rowptr, op_col, op_w, op_diag = g.operator()
    vbc = np.array([[0.3, 0.0], [0.0, -0.2], [0.1, 0.1], [0.0, 0.4]])
    b, P0, GP = g.rhs(dt, False, vbc)
    x, _ = g.cg(b, P0, rtol=1e-14, itmax=100000)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), xy=xy, dr=dr, bmin=np.array(bmin), bmax=np.array(bmax),
                        xper=xper, yper=yper, rowptr=rowptr, label=edges["label"], v1=edges["v1"], v2=edges["v2"], area=area,
                        centroid=cen, rho=rho, v=v, P=P, dt=dt, op_rowptr=op_rowptr, op_col=op_col, op_w=op_w, op_diag=op_diag,
                        vbc=vbc, b=b, GP=GP, P_solved=x)


if __name__ == "__main__":
    for nm in CASES:
        make(nm)
        print("wrote", nm)
