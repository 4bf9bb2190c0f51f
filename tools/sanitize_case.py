"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family of the library
once, at sizes that stay fast under instrumentation.  See tools/sanitize.sh."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import lvb200 as lv  # noqa: E402
from lvb200.distributed import StripGrid, StripSolver  # noqa: E402

S = lv.stepping
M = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dr = 1.0 / M
rng = np.random.default_rng(0)

# periodic jittered box: fast clip kernel, CG (fused finish) and MINRES, host-buffer paths (sync / lazy / everything lazy)
xy = lv.synthetic.jittered_lattice(M, 0)
g = lv.VoronoiGrid(lv.Rectangle((0.0, 0.0), (1.0, 1.0)), dr, xperiodic=True, yperiodic=True)
g.set_points(xy)
for lazy in (False, True, "all", "pipeline"):
    lv.remesh(g, lazy=lazy)
    lv.wait_edges(g)
# pipelined mode: deferred remesh completed by the next remesh / by find_pressure (uploads on their own stream)
v, P = lv.synthetic.taylor_green_fields(xy)
lv.remesh(g, lazy="pipeline"); lv.remesh(g, lazy="pipeline")
g.rho[...] = 1.0; g.mass[...] = 1.0 / len(xy); g.c2[...] = 100.0; g.v[...] = v; g.P[...] = P
lv.find_pressure(lv.PressureSolver(g), 0.1 * dr, 2)
lv.wait_edges(g)
lv.remesh(g)
g.rho[...] = 1.0; g.mass[...] = lv.area(g); g.c2[...] = 100.0; g.v[...] = v; g.P[...] = P
for kry in ("cg", "minres"):
    s = lv.PressureSolver(g, solver=kry)
    g.P[...] = P
    lv.find_pressure(s, 0.1 * dr, 3)
    print(kry, s.iters.tolist())

# walls + degenerate lattice: edge-list kernel (anomaly replay), per-edge wall data, stepping sweeps, projector, Lloyd
gl = (np.arange(16) + 0.5) / 16
X, Y = np.meshgrid(gl, gl, indexing="ij")
gw = lv.VoronoiGrid(lv.Rectangle((0.0, 0.0), (1.0, 1.0)), 1 / 16)
gw.set_points(np.stack([X.ravel(), Y.ravel()], 1))
lv.remesh(gw)
print("clip level / anomalies", gw.clip_info())
gw = lv.VoronoiGrid(lv.Rectangle((0.0, 0.0), (1.0, 1.0)), dr)
pts = rng.random((M * M, 2))
gw.set_points(pts)
lv.remesh(gw)
n = gw.n
gw.rho[...] = np.where(pts[:, 1] > 0.5, 1.8, 1.0); gw.mass[...] = gw.rho * lv.area(gw); gw.c2[...] = 50.0
gw.v[...] = 0.1 * rng.standard_normal((n, 2)); gw.P[...] = 1.0; gw.e[...] = 2.0; gw.mu[...] = 1e-2
gw.phase[...] = np.where(pts[:, 1] > 0.5, 0.0, 1.0); gw.quality[...] = 1.0
sw = lv.PressureSolver(gw)
lv.find_pressure(sw, 0.1 * dr, 2, boundary_velocity=lambda m, lab: np.array([np.sin(m[0]), 0.0]) if lab == -1 else np.zeros(2))
S.to_device(gw)
dt = 0.05 * dr
S.move(gw, dt); S.gravity_step(gw, (0.0, -1.0), dt); S.ideal_eos(gw, 1.4, 0.0); S.find_pressure_resident(sw, dt, 2)
S.pressure_step(gw, dt); S.find_D(gw); S.viscous_step(gw, dt, True)
S.bdary_friction(gw, dt, lambda m: np.array([1.0, 0.0]), charfun=lambda m: m[1] > 0.5)
S.find_dv(gw, dt); S.multiphase_projection(gw); S.relaxation_step(gw, dt)
S.from_device(gw)
gq = lv.VoronoiGrid(lv.Rectangle((0.0, 0.0), (1.0, 1.0)), 1 / 16)
lv.populate.populate_lloyd(gq, niterations=3, seed=2)

# strip API on one rank (library-side strip state, no peers)
sg = StripGrid(lv.Rectangle((0.0, 0.0), (1.0, 1.0)), dr, xperiodic=True, yperiodic=True, device=0)
sg.set_owned(xy, np.arange(1, len(xy) + 1))
sg.remesh(); sg.remesh()
ss = StripSolver(sg)
f = {k: torch.from_numpy(np.ascontiguousarray(a)).cuda() for k, a in
     {"mass": lv.area(g).copy(), "rho": np.ones(len(xy)), "c2": np.full(len(xy), 100.0), "P": P, "v": v}.items()}
ss.upload_fields(f["mass"], f["rho"], f["c2"], f["P"], f["v"], device=True)
print("strip", ss.find_pressure_dev(0.1 * dr, 2)[0].tolist())
sg.close()
torch.cuda.synchronize()
print("SANITIZE CASE OK")
