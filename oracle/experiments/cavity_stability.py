"""Reproduction of the open question recorded in DESIGN.md: examples/cavity.jl on the CPU restatement.

    python oracle/experiments/cavity_stability.py N [dt_factor] [t_end]

Runs the lid-driven cavity step (move!, find_pressure!, pressure_step!, find_D!, viscous_step!, bdary_friction!,
find_dv!, relaxation_step!; Re = 100, c2 = Inf, dt = dt_factor * min(0.1 dr, 0.1 Re dr^2), random + 100 Lloyd
iterations as populate_lloyd!) and reports when the velocity leaves the physical range, together with the drift of
mass/area that precedes it.  TEST INFRASTRUCTURE (uses the oracle); not part of the product or of the test-suite.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as orc  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dtf = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    t_end = float(sys.argv[3]) if len(sys.argv) > 3 else 3.0
    Re, dr = 100.0, 1.0 / N
    dt = dtf * min(0.1 * dr, 0.1 * Re * dr * dr)                       # cavity.jl:31
    og = orc.OracleGrid((0.0, 0.0), (1.0, 1.0), dr)
    rng = np.random.default_rng(1)
    n = round(1.0 / dr ** 2)
    og.set_points(rng.random((n, 2)))                                  # populate_rand!  populate.jl:75-93
    for _ in range(100):                                               # populate_lloyd!  populate.jl:132-145
        assert og.remesh() == 0
        og.set("x", og.centroid())
    assert og.remesh() == 0
    og.set("rho", np.ones(n)); og.set("mass", og.area().copy()); og.set("mu", np.full(n, 1.0 / Re)); og.set("c2", np.full(n, np.inf))
    og.set("v", np.zeros((n, 2))); og.set("P", np.zeros(n)); og.set("e", np.zeros(n))
    lid = np.array([[1.0, 0.0], [0.0, 0.0], [0.0, 0.0], [0.0, 0.0]])   # vDirichlet  cavity.jl:41-44
    t, k = 0.0, 0
    while t < t_end:
        if og.move(dt) != 0:
            print(f"move! failed at step {k}, t = {t:.4f}"); return
        og.find_pressure(dt, 10, solver="minres")
        og.pressure_step(dt)
        og.find_D(); og.viscous_step(dt, False)
        og.bdary_friction(dt, lid)
        og.find_dv(dt, 1.0)
        if og.relaxation_step(dt, True) != 0:
            print(f"relaxation_step! failed at step {k}, t = {t:.4f}"); return
        t += dt; k += 1
        v = og.get("v")
        if k % 50 == 0 or not np.isfinite(v).all() or np.abs(v).max() > 3.0:
            r = og.get("mass") / og.area()
            print(f"step {k} t = {t:.4f} |v|max = {np.abs(v).max():.4g} mass/area in [{r.min():.3f}, {r.max():.3f}] "
                  f"KE = {0.5 * (og.get('mass') * (v ** 2).sum(1)).sum():.5f}", flush=True)
        if not np.isfinite(v).all() or np.abs(v).max() > 3.0:
            print("unstable"); return
    print("stable to t =", t)


if __name__ == "__main__":
    main()
