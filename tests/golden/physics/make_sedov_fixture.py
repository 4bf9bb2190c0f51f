"""Down-samples the reference's semi-analytic Sedov profile (examples/reference/sedov.csv: r, rho, v, P at t = 1 for
E = 0.3, rho0 = 1, gamma = 1.4 -- the curve examples/sedov.jl plots its result against) to a small fixture.

    python tests/golden/physics/make_sedov_fixture.py        # needs /root/reference; writes tests/golden/physics/sedov_profile.npz
"""
import os

import numpy as np

SRC = "/root/reference/examples/reference/sedov.csv"
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    ref = np.loadtxt(SRC, delimiter=",", skiprows=1)
    r = np.linspace(ref[0, 0], ref[-1, 0], 400)
    out = {"r": r, "rho": np.interp(r, ref[:, 0], ref[:, 1]), "v": np.interp(r, ref[:, 0], ref[:, 2]), "P": np.interp(r, ref[:, 0], ref[:, 3]),
           "r_shock": ref[np.argmax(ref[:, 1]), 0], "source": SRC + " (columns r, rho, v, P)"}
    np.savez_compressed(os.path.join(HERE, "sedov_profile.npz"), **out)
    print("shock radius", out["r_shock"], "peak density", out["rho"].max())
