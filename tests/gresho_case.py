"""examples/gresho.jl (BASELINE config 0) restated for the tests: seeding, exact solution, error norm."""
import numpy as np

RHO0, GAMMA, MA = 1.0, 1.4, 0.1                                            # gresho.jl:24-37
C0 = 1.0 / MA
P0 = RHO0 * C0 ** 2 / GAMMA
BMIN, BMAX = (-0.5, -0.5), (0.5, 0.5)


def circ_points(dr, center=(0.0, 0.0)):
    """populate_circ!  populate.jl:18-35 (generators on concentric circles, those inside the rectangle)"""
    c = np.asarray(center, dtype=np.float64)
    r_max = max(np.hypot(px - c[0], py - c[1]) for px in (BMIN[0], BMAX[0]) for py in (BMIN[1], BMAX[1]))
    pts, r = [], 0.5 * dr
    while r <= r_max:
        k_max = int(round(2.0 * np.pi * r / dr))
        th = 2.0 * np.pi * np.arange(1, k_max + 1) / k_max
        pts.append(np.stack([c[0] + r * np.cos(th), c[1] + r * np.sin(th)], 1))
        r += dr
    p = np.concatenate(pts)
    return p[(p[:, 0] >= BMIN[0]) & (p[:, 0] <= BMAX[0]) & (p[:, 1] >= BMIN[1]) & (p[:, 1] <= BMAX[1])]


def v_exact(x):                                                             # gresho.jl:43-50
    r = np.sqrt((x ** 2).sum(1))
    om = np.where(r < 0.2, 5.0, np.where(r < 0.4, 2.0 / np.maximum(r, 1e-300) - 5.0, 0.0))
    return om[:, None] * np.stack([-x[:, 1], x[:, 0]], 1)


def P_exact(x, Pmin=P0):                                                    # gresho.jl:52-59 (ideal gas branch: Pmin = P0)
    r = np.sqrt((x ** 2).sum(1))
    rr = np.maximum(r, 1e-300)
    return np.where(r < 0.2, Pmin + 12.5 * r ** 2,
                    np.where(r < 0.4, Pmin + 4.0 + 4 * np.log(5 * rr) - 20.0 * r + 12.5 * r ** 2, Pmin - 2.0 + 4 * np.log(2)))


def initial_fields(x, area):                                                # ic!  gresho.jl:64-70
    v, P = v_exact(x), P_exact(x)
    return {"v": v, "rho": np.full(len(x), RHO0), "mass": RHO0 * area, "P": P, "e": 0.5 * (v ** 2).sum(1) + P / (RHO0 * (GAMMA - 1.0)),
            "mu": np.zeros(len(x))}


def l2_error(x, v, mass):                                                   # postproc!  gresho.jl:121-126, relative to |v_exact|
    ve = v_exact(x)
    return float(np.sqrt((mass * ((v - ve) ** 2).sum(1)).sum()) / np.sqrt((mass * (ve ** 2).sum(1)).sum()))
