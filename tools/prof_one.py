"""Small driver for ncu captures: one remesh + find_pressure at side M (default 1024) on cuda:0.
Usage (under gpurun):
  ncu --set full --import-source on --clock-control none -k regex:'k_clip_fast|k_assemble|k_rhs|k_matvec' -c 12 \
      -o gpurun_out/cap python tools/prof_one.py 1024
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import lvb200 as lv  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
niter = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dr = 1.0 / M
xy = lv.synthetic.jittered_lattice(M, 0)
g = lv.VoronoiGrid(lv.Rectangle((0.0, 0.0), (1.0, 1.0)), dr, xperiodic=True, yperiodic=True, device=0)
g.set_points(xy)
xy_dev = torch.from_numpy(xy).cuda()
g.remesh_dev(xy_dev)
g.remesh_dev(xy_dev)
s = lv.PressureSolver(g)
_, _, area, _ = g.mesh_download(M * M, edges=False)
v, P = lv.synthetic.taylor_green_fields(xy)
f = {k: torch.from_numpy(np.ascontiguousarray(a)).cuda() for k, a in
     {"mass": area.copy(), "rho": np.ones(M * M), "c2": np.full(M * M, 100.0), "P": P, "v": v}.items()}
s.upload_fields(f["mass"], f["rho"], f["c2"], f["P"], f["v"], device=True)
it, _ = s.find_pressure_dev(0.1 * dr, niter)
torch.cuda.synchronize()
print("iters", it.tolist(), "launches", g.launch_count())
