"""examples/sedov.jl restated as a driver over any object with the sweep methods (the oracle grid on CPU, the
device-resident stepping API on the GPU), plus the comparison with the reference's own semi-analytic profile
(examples/reference/sedov.csv, frozen in tests/golden/physics/sedov_profile.npz by make_sedov_fixture.py)."""
import os

import numpy as np

RHO0, GAMMA, P0, R_BOMB, E_BOMB, CFL = 1.0, 1.4, 1e-8, 0.05, 0.3, 0.1      # sedov.jl:38-52
T_BOMB = np.sqrt(RHO0 / E_BOMB * R_BOMB ** 5)


def initial_fields(x, area):
    """ic! + detonate_bomb!  sedov.jl:54-80"""
    n = len(x)
    P = np.full(n, P0)
    inb = np.sqrt((x ** 2).sum(1)) < R_BOMB
    P[inb] = (GAMMA - 1.0) * E_BOMB / area[inb].sum()
    return {"rho": np.full(n, RHO0), "mass": RHO0 * area, "v": np.zeros((n, 2)), "P": P, "e": P / (RHO0 * (GAMMA - 1.0)),
            "mu": np.zeros(n)}


def time_steps(dr, t_end=1.0):
    """the adaptive dt of step!  sedov.jl:107-109"""
    t = T_BOMB
    while t < t_end:
        v_shock = 0.4 * t ** (-0.6) * (E_BOMB / RHO0) ** 0.2
        dt = CFL * dr / (np.sqrt(6.0) * v_shock)
        yield dt
        t += dt


def compare_with_reference(x, rho, dr):
    """Radially binned density against the semi-analytic solution at t = 1.  Returns a dict of the quantities the tests
    bound."""
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "physics", "sedov_profile.npz"))
    r = np.sqrt((x ** 2).sum(1))
    bins = np.arange(0.0, 1.0 + 1e-12, dr)
    idx = np.digitize(r, bins)
    rc = 0.5 * (bins[1:] + bins[:-1])
    prof = np.array([rho[idx == i].mean() if (idx == i).any() else np.nan for i in range(1, len(bins))])
    refp = np.interp(rc, ref["r"], ref["rho"], left=float(ref["rho"][0]), right=RHO0)
    refp[rc > float(ref["r_shock"])] = RHO0
    wake = (rc > 0.40) & (rc < float(ref["r_shock"]) - 3 * dr) & np.isfinite(prof)
    ahead = (rc > float(ref["r_shock"]) + 5 * dr) & np.isfinite(prof)
    return {"r_shock_ref": float(ref["r_shock"]), "r_shock": float(rc[np.nanargmax(prof)]), "peak": float(np.nanmax(prof)),
            "wake_rel_err": float(np.mean(np.abs(prof[wake] - refp[wake]) / refp[wake])),
            "ahead_err": float(np.abs(prof[ahead] - RHO0).max())}
