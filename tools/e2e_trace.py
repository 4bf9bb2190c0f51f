"""Timeline of the pipelined host-buffer step (LV_PIPE_TRACE=1): python tools/e2e_trace.py [M]"""
import os, sys, time
os.environ["LV_PIPE_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import lvb200 as lv
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
os.system("lscpu | egrep 'Model name|^CPU\\(s\\)|Socket|NUMA node\\(s\\)|Thread' >&2; nproc >&2; free -g | head -2 >&2")
dr = 1.0 / M; n = M * M
xy = lv.synthetic.jittered_lattice(M, 0)
g = lv.VoronoiGrid(lv.Rectangle((0.0, 0.0), (1.0, 1.0)), dr, xperiodic=True, yperiodic=True, device=0)
g.set_points(xy)
lv.remesh(g, edges=False)
v, P = lv.synthetic.taylor_green_fields(xy)
for name, val in (("rho", 1.0), ("mass", lv.area(g)), ("c2", 100.0), ("v", v), ("P", P)):
    getattr(g, name)[...] = val
s = lv.PressureSolver(g, solver="pcg")
for it in range(4):
    g.P[...] = P
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    print(f"--- step {it}", file=sys.stderr)
    lv.remesh(g, lazy="pipeline"); t1 = time.perf_counter()
    lv.remesh(g, lazy="pipeline"); t2 = time.perf_counter()
    lv.find_pressure(s, 0.1 * dr, 10); t3 = time.perf_counter()
    lv.wait_edges(g); t4 = time.perf_counter()
    print(f"step {it}: remesh {1e3*(t1-t0):.1f} remesh {1e3*(t2-t1):.1f} find_pressure {1e3*(t3-t2):.1f} wait {1e3*(t4-t3):.1f} total {1e3*(t4-t0):.1f} ms", file=sys.stderr)
