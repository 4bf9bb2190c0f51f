"""VTK / pvd / csv output fed from the flat mesh arrays.

Mirrors IO.jl:14-72 and the output part of simulation.jl:35-94: ``export_grid`` writes the Voronoi polygons as
PolyData polys (the chain of ``v1`` vertices of every polygon, IO.jl:18-26) with per-polygon cell data,
``export_points`` the generators as a vertex cloud, ``run`` is the fixed-dt loop with frame cadence, ``cells.pvd`` /
``points.pvd`` collections and ``simdata.csv``.  Files are plain (ascii) VTK XML that ParaView reads; the reference
writes the same data model through WriteVTK.jl (compressed binary).
"""
from __future__ import annotations

import os

import numpy as np


def _data_array(name, arr, ncomp=1, dtype="Float64"):
    a = np.asarray(arr)
    fmt = "%.17g" if dtype.startswith("Float") else "%d"
    body = "\n".join(" ".join(fmt % v for v in row) for row in a.reshape(-1, ncomp if ncomp > 1 else 1))
    nc = f' NumberOfComponents="{ncomp}"' if ncomp > 1 else ""
    return f'<DataArray type="{dtype}" Name="{name}"{nc} format="ascii">\n{body}\n</DataArray>\n'


def _vec3(a):
    a = np.asarray(a, dtype=np.float64).reshape(-1, 2)
    return np.concatenate([a, np.zeros((a.shape[0], 1))], axis=1)  # Vec3  IO.jl:1-3


def _datasets(grid, names):
    out = ""
    for nm in names:
        if not hasattr(grid, nm):
            raise ValueError(f"Cannot export variable {nm} because it does not exist.")  # IO.jl:65-67
        a = np.asarray(getattr(grid, nm))
        out += _data_array(nm, _vec3(a), 3) if a.ndim == 2 else _data_array(nm, a)
    return out


def export_grid(grid, filename: str, *variables) -> str:
    """export_grid(grid, filename, vars...)  IO.jl:14-33"""
    if not filename.endswith(".vtp"):
        filename += ".vtp"
    rowptr, edges = grid.rowptr, grid.edges
    pts = _vec3(edges["v1"])                                   # one vertex per edge, in chain order
    n, m = len(rowptr) - 1, pts.shape[0]
    xml = ('<?xml version="1.0"?>\n<VTKFile type="PolyData" version="1.0" byte_order="LittleEndian">\n<PolyData>\n'
           f'<Piece NumberOfPoints="{m}" NumberOfVerts="0" NumberOfLines="0" NumberOfStrips="0" NumberOfPolys="{n}">\n'
           "<Points>\n" + _data_array("Points", pts, 3) + "</Points>\n<Polys>\n"
           + _data_array("connectivity", np.arange(m), 1, "Int64") + _data_array("offsets", rowptr[1:], 1, "Int64")
           + "</Polys>\n<CellData>\n" + _datasets(grid, variables) + "</CellData>\n</Piece>\n</PolyData>\n</VTKFile>\n")
    with open(filename, "w") as f:
        f.write(xml)
    return filename


def export_points(grid, filename: str, *variables) -> str:
    """export_points(grid, filename, vars...)  IO.jl:50-57"""
    if not filename.endswith(".vtp"):
        filename += ".vtp"
    pts = _vec3(grid.x)
    n = pts.shape[0]
    xml = ('<?xml version="1.0"?>\n<VTKFile type="PolyData" version="1.0" byte_order="LittleEndian">\n<PolyData>\n'
           f'<Piece NumberOfPoints="{n}" NumberOfVerts="{n}" NumberOfLines="0" NumberOfStrips="0" NumberOfPolys="0">\n'
           "<Points>\n" + _data_array("Points", pts, 3) + "</Points>\n<Verts>\n"
           + _data_array("connectivity", np.arange(n), 1, "Int64") + _data_array("offsets", np.arange(1, n + 1), 1, "Int64")
           + "</Verts>\n<PointData>\n" + _datasets(grid, variables) + "</PointData>\n</Piece>\n</PolyData>\n</VTKFile>\n")
    with open(filename, "w") as f:
        f.write(xml)
    return filename


class _Pvd:
    """paraview_collection  (simulation.jl:49-50, 80-91)"""

    def __init__(self, path):
        self.path, self.entries = path, []

    def add(self, t, file):
        self.entries.append((t, os.path.basename(file)))

    def save(self):
        with open(self.path, "w") as f:
            f.write('<?xml version="1.0"?>\n<VTKFile type="Collection" version="1.0" byte_order="LittleEndian">\n<Collection>\n')
            for t, fn in self.entries:
                f.write(f'<DataSet timestep="{t:.17g}" part="0" file="{fn}"/>\n')
            f.write("</Collection>\n</VTKFile>\n")


def run(sim, dt: float, t_end: float, step, nframes: int = 100, path: str = "results", save_points: bool = False,
        save_grid: bool = True, save_csv: bool = True, vtp_vars=(), csv_vars=(), postproc=None, sync=None):
    """run!(sim, dt, t_end, step!; ...)  simulation.jl:35-94.  ``sim.grid`` is the VoronoiGrid; ``sync(sim)`` (optional) brings
    device-resident state to the host right before a frame is written."""
    os.makedirs(path, exist_ok=True)
    pvd_c, pvd_p = _Pvd(os.path.join(path, "cells.pvd")), _Pvd(os.path.join(path, "points.pvd"))
    csv = {"time": []}
    for var in csv_vars:
        if var == "time":
            raise ValueError("csv_vars cannot be :time")
        csv[var] = []
    k = nframe = 0
    k_frame = max(int(round(t_end / (nframes * dt))), 1) if nframes > 0 else 2 ** 62
    t = 0.0
    while t < t_end:
        k += 1
        step(sim, t)
        if k % k_frame == 0:
            if sync is not None:
                sync(sim)
            if postproc is not None:
                postproc(sim, t)
            csv["time"].append(t)
            for var in csv_vars:
                csv[var].append(float(getattr(sim, var)))
            if save_grid:
                pvd_c.add(t, export_grid(sim.grid, os.path.join(path, f"cframe{nframe}.vtp"), *vtp_vars))
            if save_points:
                pvd_p.add(t, export_points(sim.grid, os.path.join(path, f"pframe{nframe}.vtp"), *vtp_vars))
            nframe += 1
        t += dt
    if save_grid:
        pvd_c.save()
    if save_points:
        pvd_p.save()
    if save_csv:
        keys = list(csv)
        with open(os.path.join(path, "simdata.csv"), "w") as f:
            f.write(",".join(keys) + "\n")
            for row in zip(*[csv[kk] for kk in keys]):
                f.write(",".join("%.17g" % v for v in row) + "\n")
