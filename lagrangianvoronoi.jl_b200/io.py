"""VTK / pvd / csv output fed from the flat mesh arrays.

Mirrors IO.jl:14-72 and the output part of simulation.jl:35-94: ``export_grid`` writes the Voronoi polygons as
PolyData polys (the chain of ``v1`` vertices of every polygon, IO.jl:18-26) with per-polygon cell data,
``export_points`` the generators as a vertex cloud, ``run`` is the fixed-dt loop with frame cadence, ``cells.pvd`` /
``points.pvd`` collections and ``simdata.csv``.  Files are VTK XML PolyData with appended raw binary data (optionally zlib
compressed) -- the container WriteVTK.jl produces for the reference -- written straight from the numpy buffers.
"""
from __future__ import annotations

import os

import numpy as np


_VTK_TYPES = {np.dtype("float64"): "Float64", np.dtype("int64"): "Int64", np.dtype("int32"): "Int32", np.dtype("float32"): "Float32"}


class _VtpWriter:
    """VTK XML PolyData with APPENDED RAW data (what WriteVTK.jl's ``vtk_grid`` writes by default: ``append = true``), optionally
    zlib-compressed per block like WriteVTK's ``compress = true``.  Arrays are written straight from their numpy buffers --
    no per-number formatting -- so a 16M-cell frame costs a few seconds of memcpy / file I/O instead of 10^8 ``%`` operations.
    ``ascii=True`` keeps the old plain-text form (small files, debugging)."""

    BLOCK = 1 << 22  # uncompressed bytes per zlib block

    def __init__(self, ascii: bool = False, compress: bool = False):
        self.ascii, self.compress = ascii, compress
        self.blobs, self.offset = [], 0

    def array(self, name, arr, ncomp=1):
        a = np.ascontiguousarray(arr)
        if a.dtype not in _VTK_TYPES:
            a = a.astype(np.float64 if a.dtype.kind == "f" else np.int64)
        typ = _VTK_TYPES[a.dtype]
        nc = f' NumberOfComponents="{ncomp}"' if ncomp > 1 else ""
        if self.ascii:
            flat = a.reshape(-1, ncomp if ncomp > 1 else 1)
            fmt = "%.17g" if a.dtype.kind == "f" else "%d"
            body = "\n".join(" ".join(fmt % v for v in row) for row in flat)
            return f'<DataArray type="{typ}" Name="{name}"{nc} format="ascii">\n{body}\n</DataArray>\n'
        raw = memoryview(a).cast("B")
        if self.compress:
            import zlib
            nb = max(1, (len(raw) + self.BLOCK - 1) // self.BLOCK)
            parts = [zlib.compress(raw[k * self.BLOCK:(k + 1) * self.BLOCK], 1) for k in range(nb)]
            last = len(raw) - (nb - 1) * self.BLOCK
            head = np.array([nb, self.BLOCK, 0 if last == self.BLOCK else last] + [len(q) for q in parts], dtype="<u8").tobytes()
            chunk = [head] + parts
        else:
            chunk = [np.array([len(raw)], dtype="<u8").tobytes(), raw]
        off = self.offset
        self.blobs += chunk
        self.offset += sum(len(q) for q in chunk)
        return f'<DataArray type="{typ}" Name="{name}"{nc} format="appended" offset="{off}"/>\n'

    def header(self):
        comp = ' compressor="vtkZLibDataCompressor"' if (self.compress and not self.ascii) else ""
        return f'<?xml version="1.0"?>\n<VTKFile type="PolyData" version="1.0" byte_order="LittleEndian" header_type="UInt64"{comp}>\n'

    def write(self, filename, body):
        with open(filename, "wb") as f:
            f.write((self.header() + body).encode())
            if not self.ascii:
                f.write(b'<AppendedData encoding="raw">\n_')
                for q in self.blobs:
                    f.write(q)
                f.write(b"\n</AppendedData>\n")
            f.write(b"</VTKFile>\n")


def _vec3(a):
    a = np.asarray(a, dtype=np.float64).reshape(-1, 2)
    out = np.zeros((a.shape[0], 3))
    out[:, :2] = a
    return out  # Vec3  IO.jl:1-3


def _datasets(w, grid, names):
    out = ""
    for nm in names:
        if not hasattr(grid, nm):
            raise ValueError(f"Cannot export variable {nm} because it does not exist.")  # IO.jl:65-67
        a = np.asarray(getattr(grid, nm))
        out += w.array(nm, _vec3(a), 3) if a.ndim == 2 else w.array(nm, a.astype(np.float64, copy=False))
    return out


def export_grid(grid, filename: str, *variables, ascii: bool = False, compress: bool = False) -> str:
    """export_grid(grid, filename, vars...)  IO.jl:14-33"""
    if not filename.endswith(".vtp"):
        filename += ".vtp"
    rowptr, edges = grid.rowptr, grid.edges
    pts = _vec3(edges["v1"])                                   # one vertex per edge, in chain order
    n, m = len(rowptr) - 1, pts.shape[0]
    w = _VtpWriter(ascii, compress)
    body = ("<PolyData>\n"
            f'<Piece NumberOfPoints="{m}" NumberOfVerts="0" NumberOfLines="0" NumberOfStrips="0" NumberOfPolys="{n}">\n'
            "<Points>\n" + w.array("Points", pts, 3) + "</Points>\n<Polys>\n"
            + w.array("connectivity", np.arange(m, dtype=np.int64)) + w.array("offsets", np.asarray(rowptr[1:], dtype=np.int64))
            + "</Polys>\n<CellData>\n" + _datasets(w, grid, variables) + "</CellData>\n</Piece>\n</PolyData>\n")
    w.write(filename, body)
    return filename


def export_points(grid, filename: str, *variables, ascii: bool = False, compress: bool = False) -> str:
    """export_points(grid, filename, vars...)  IO.jl:50-57"""
    if not filename.endswith(".vtp"):
        filename += ".vtp"
    pts = _vec3(grid.x)
    n = pts.shape[0]
    w = _VtpWriter(ascii, compress)
    body = ("<PolyData>\n"
            f'<Piece NumberOfPoints="{n}" NumberOfVerts="{n}" NumberOfLines="0" NumberOfStrips="0" NumberOfPolys="0">\n'
            "<Points>\n" + w.array("Points", pts, 3) + "</Points>\n<Verts>\n"
            + w.array("connectivity", np.arange(n, dtype=np.int64)) + w.array("offsets", np.arange(1, n + 1, dtype=np.int64))
            + "</Verts>\n<PointData>\n" + _datasets(w, grid, variables) + "</PointData>\n</Piece>\n</PolyData>\n")
    w.write(filename, body)
    return filename


def read_vtp(filename: str) -> dict:
    """Minimal reader of the files written above (tests, post-processing): {array name: numpy array} plus the piece counts."""
    import re
    import zlib
    blob = open(filename, "rb").read()
    k = blob.find(b'<AppendedData encoding="raw">')
    xml = blob if k < 0 else blob[:k]
    text = xml.decode()
    data0 = None if k < 0 else blob.index(b"_", k) + 1
    compressed = "vtkZLibDataCompressor" in text
    out = {"_counts": {m.group(1): int(m.group(2)) for m in re.finditer(r'NumberOf(\w+)="(\d+)"', text)}}
    types = {v: k2 for k2, v in _VTK_TYPES.items()}
    for m in re.finditer(r'<DataArray type="(\w+)" Name="(\w+)"(?: NumberOfComponents="(\d+)")? format="(\w+)"(?: offset="(\d+)")?\s*(/>|>)', text):
        typ, name, nc, fmt, off = m.group(1), m.group(2), int(m.group(3) or 1), m.group(4), m.group(5)
        dt = types[typ]
        if fmt == "ascii":
            end = text.index("</DataArray>", m.end())
            a = np.array(text[m.end():end].split(), dtype=dt)
        else:
            p = data0 + int(off)
            if compressed:
                nb, bs, last = (int(v) for v in np.frombuffer(blob, "<u8", 3, p))
                sizes = np.frombuffer(blob, "<u8", nb, p + 24)
                q = p + 24 + 8 * nb
                parts = []
                for sz in sizes:
                    parts.append(zlib.decompress(blob[q:q + int(sz)]))
                    q += int(sz)
                a = np.frombuffer(b"".join(parts), dtype=dt)
            else:
                nbytes = int(np.frombuffer(blob, "<u8", 1, p)[0])
                a = np.frombuffer(blob, dtype=dt, count=nbytes // dt.itemsize, offset=p + 8)
        out[name] = a.reshape(-1, nc) if nc > 1 else a
    return out


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
