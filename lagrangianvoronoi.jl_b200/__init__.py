"""B200-native (sm_100a) mesh-and-pressure hot path of LagrangianVoronoi.jl.

Host-side mirror of the reference's Julia call surface for this path
(``VoronoiGrid``, ``remesh!``, ``PressureSolver``, ``find_pressure!``, ``mul!``) on top of
the C-ABI shared library ``lib/liblvb200.so`` (include/lv_capi.h).  There is no CPU
fallback: every compute call goes to the CUDA library and fails loudly without it.
"""
from ._capi import LvError, build_library, library_path, load_library, EDGE_DTYPE  # noqa: F401
from .host import (Rectangle, VoronoiGrid, PressureSolver, remesh, find_pressure, area, centroid,  # noqa: F401
                   neighbors_csr, mul, wait_edges)
from . import synthetic  # noqa: F401
from . import distributed  # noqa: F401
from . import stepping  # noqa: F401
from . import populate  # noqa: F401
from . import io  # noqa: F401

__all__ = ["LvError", "build_library", "library_path", "load_library", "EDGE_DTYPE", "Rectangle", "VoronoiGrid",
           "PressureSolver", "remesh", "find_pressure", "area", "centroid", "neighbors_csr", "mul", "wait_edges", "synthetic", "distributed", "stepping", "populate", "io"]
