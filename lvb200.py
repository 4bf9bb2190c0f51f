"""Import alias for the package directory ``lagrangianvoronoi.jl_b200``.

The directory name contains a dot, so ``import lagrangianvoronoi.jl_b200`` cannot work;
``import lvb200`` loads that directory as a regular package under the name ``lvb200``.
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lagrangianvoronoi.jl_b200")
_spec = importlib.util.spec_from_file_location("lvb200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["lvb200"] = _mod
_spec.loader.exec_module(_mod)
