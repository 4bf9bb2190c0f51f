"""ctypes binding of include/lv_capi.h (the same symbols the Julia shim ``ccall``s)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "lib", "liblvb200.so")

EDGE_DTYPE = np.dtype([("v1", "<f8", (2,)), ("v2", "<f8", (2,)), ("label", "<i8")])  # geometry.jl:82-87
assert EDGE_DTYPE.itemsize == 40

LV_OK, LV_EINVAL, LV_EDESTROYED, LV_ENAN, LV_ECUDA, LV_ECAPACITY = range(6)
LV_SOLVER_CG, LV_SOLVER_MINRES, LV_SOLVER_PCG = 0, 1, 2


def solver_kind(name: str) -> int:
    """"cg" (north star), "minres" (the reference's Krylov method) or "pcg" (CG with the Jacobi preconditioner 1/A_ii)."""
    try:
        return {"cg": LV_SOLVER_CG, "minres": LV_SOLVER_MINRES, "pcg": LV_SOLVER_PCG}[name]
    except KeyError:
        raise ValueError(f"unknown Krylov method {name!r}") from None
PROF_SLOTS = {"cells": 0, "clip": 1, "assemble": 2, "matvec": 3, "vecops": 4}

# every symbol include/lv_capi.h declares; tests check the library exports all of them
SYMBOLS = [
    "lv_create", "lv_destroy", "lv_last_error", "lv_set_rects", "lv_grid_info", "lv_magic_path", "lv_set_stream",
    "lv_sync", "lv_remesh", "lv_remesh_dev", "lv_mesh_nnz", "lv_mesh_download", "lv_mesh_faces", "lv_mesh_hash", "lv_boundary_edges", "lv_set_boundary_velocity", "lv_clip_info", "lv_set_async_edges", "lv_mesh_wait", "lv_wire_expand",
    "lv_pressure_create", "lv_pressure_destroy", "lv_fields_upload", "lv_fields_upload_dev", "lv_pressure_download",
    "lv_pressure_assemble", "lv_pressure_operator", "lv_pressure_matvec", "lv_pressure_rhs", "lv_find_pressure",
    "lv_find_pressure_dev", "lv_pressure_solve", "lv_prof_enable", "lv_prof_reset", "lv_prof_get",
    "lv_launch_count", "lv_device_bytes", "lv_comm_unique_id", "lv_comm_init", "lv_remesh_owned_dev", "lv_device_array",
    "lv_halo_plan", "lv_halo_exchange_dev", "lv_strip_setup", "lv_strip_map", "lv_strip_set_owned", "lv_strip_remesh", "lv_strip_set_rows", "lv_state_attach_strip", "lv_mailbox_export", "lv_mailbox_plan", "lv_peer_disable", "lv_peer_close",
    "lv_state_set", "lv_state_get", "lv_state_ptr", "lv_state_remesh", "lv_step_move", "lv_step_eos", "lv_step_find_pressure",
    "lv_step_pressure_step", "lv_step_gravity", "lv_step_find_D", "lv_step_viscous_step", "lv_step_bdary_friction", "lv_step_bdary_friction_ex", "lv_step_find_dv", "lv_step_relaxation_step", "lv_step_lloyd", "lv_step_multiphase_projection", "lv_step_multiphase_apply",
]


class LvError(RuntimeError):
    """Raised for every non-zero status of the C ABI; mirrors the reference's exceptions."""

    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


class GridDesc(C.Structure):
    _fields_ = [("dr", C.c_double), ("h", C.c_double), ("r_max", C.c_double), ("xperiodic", C.c_int32),
                ("yperiodic", C.c_int32), ("bmin", C.c_double * 2), ("bmax", C.c_double * 2)]


def library_path() -> str:
    return _LIB


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into lib/liblvb200.so (nvcc cross-compiles without a GPU)."""
    src_dir = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(_HERE, "..", "include", "lv_capi.h"))
    stale = (not os.path.exists(_LIB)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs)
    if force or stale:
        cmd = ["make", "-C", src_dir] + ([] if verbose else ["-s"])
        subprocess.run(cmd, check=True)
    return _LIB


_lib = None


def load_library() -> C.CDLL:
    """Load liblvb200.so.  No fallback: a missing library is a hard error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise RuntimeError(f"{_LIB} is missing: run __graft_entry__.build() (nvcc, sm_100a). "
                           "There is no CPU fallback for this path.")
    L = C.CDLL(_LIB)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_void_p
    i32p = C.POINTER(C.c_int32)
    L.lv_create.argtypes = [C.POINTER(GridDesc), C.c_int32, C.POINTER(vp)]
    L.lv_destroy.argtypes = [vp]
    L.lv_last_error.argtypes = [vp]
    L.lv_last_error.restype = C.c_char_p
    L.lv_set_rects.argtypes = [vp, dp, dp, dp, dp]
    L.lv_grid_info.argtypes = [vp, ip, ip, dp, ip]
    L.lv_magic_path.argtypes = [vp, C.c_int64, ip, ip, dp, ip]
    L.lv_set_stream.argtypes = [vp, vp, C.c_int32]
    L.lv_sync.argtypes = [vp]
    L.lv_remesh.argtypes = [vp, C.c_int64, vp, vp, vp, C.c_int64, ip, vp, vp]
    L.lv_remesh_dev.argtypes = [vp, C.c_int64, vp]
    L.lv_mesh_nnz.argtypes = [vp, ip]
    L.lv_mesh_download.argtypes = [vp, vp, vp, C.c_int64, vp, vp]
    L.lv_mesh_faces.argtypes = [vp, vp, vp, C.c_int64]
    L.lv_clip_info.argtypes = [vp, i32p, ip]
    L.lv_set_async_edges.argtypes = [vp, C.c_int32]
    L.lv_mesh_wait.argtypes = [vp]
    L.lv_wire_expand.argtypes = [vp, vp, C.c_int64, C.c_int32, vp, vp]
    L.lv_pressure_create.argtypes = [vp]
    L.lv_pressure_destroy.argtypes = [vp]
    L.lv_fields_upload.argtypes = [vp, vp, vp, vp, vp, vp]
    L.lv_fields_upload_dev.argtypes = [vp, vp, vp, vp, vp, vp]
    L.lv_pressure_download.argtypes = [vp, vp]
    L.lv_pressure_assemble.argtypes = [vp, C.c_double]
    L.lv_pressure_operator.argtypes = [vp, vp, vp, vp, C.c_int64, vp]
    L.lv_pressure_matvec.argtypes = [vp, vp, vp]
    L.lv_pressure_rhs.argtypes = [vp, C.c_double, C.c_int32, vp, vp, vp]
    L.lv_find_pressure.argtypes = [vp, C.c_double, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32,
                                   vp, vp, vp, vp, vp, vp, vp, C.c_int64, vp, i32p, dp]
    L.lv_boundary_edges.argtypes = [vp, ip, vp, vp, vp, C.c_int64]
    L.lv_set_boundary_velocity.argtypes = [vp, vp, C.c_int64]
    L.lv_step_bdary_friction_ex.argtypes = [vp, C.c_double, vp, vp, vp, vp, C.c_int64]
    L.lv_find_pressure_dev.argtypes = [vp, C.c_double, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32,
                                       vp, i32p, dp]
    L.lv_pressure_solve.argtypes = [vp, C.c_int32, vp, vp, C.c_double, C.c_double, C.c_int32, i32p, dp]
    L.lv_comm_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    L.lv_comm_init.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(C.c_uint8)]
    L.lv_remesh_owned_dev.argtypes = [vp, C.c_int64, vp, vp, vp]
    L.lv_device_array.argtypes = [vp, C.c_int32, C.POINTER(vp), ip]
    L.lv_halo_plan.argtypes = [vp, C.c_int32, i32p, ip, vp, ip, vp]
    L.lv_halo_exchange_dev.argtypes = [vp, vp, C.c_int32]
    L.lv_strip_setup.argtypes = [vp, C.c_int32, i32p, i32p, i32p, i32p, C.c_int64, C.c_int64, C.POINTER(C.c_uint8)]
    L.lv_strip_map.argtypes = [vp, C.POINTER(C.c_uint8)]
    L.lv_strip_set_owned.argtypes = [vp, C.c_int64, vp, C.c_int32, vp]
    L.lv_strip_remesh.argtypes = [vp, ip]
    L.lv_strip_set_rows.argtypes = [vp, C.c_int32, i32p]
    L.lv_state_attach_strip.argtypes = [vp]
    L.lv_peer_close.argtypes = [vp]
    L.lv_mesh_hash.argtypes = [vp, vp, C.POINTER(C.c_uint64)]
    L.lv_step_multiphase_apply.argtypes = [vp, vp, vp, vp]
    L.lv_mailbox_export.argtypes = [vp, C.POINTER(C.c_uint8)]
    L.lv_mailbox_plan.argtypes = [vp, C.c_int32, C.POINTER(C.c_uint8)]
    L.lv_peer_disable.argtypes = [vp]
    L.lv_state_set.argtypes = [vp, C.c_char_p, vp, C.c_int64]
    L.lv_state_get.argtypes = [vp, C.c_char_p, vp]
    L.lv_state_ptr.argtypes = [vp, C.c_char_p, C.POINTER(vp), ip]
    L.lv_state_remesh.argtypes = [vp]
    L.lv_step_move.argtypes = [vp, C.c_double]
    L.lv_step_eos.argtypes = [vp, C.c_double, C.c_double, C.c_int32]
    L.lv_step_find_pressure.argtypes = [vp, C.c_double, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32, vp, i32p, dp]
    L.lv_step_pressure_step.argtypes = [vp, C.c_double]
    L.lv_step_gravity.argtypes = [vp, C.c_double, C.c_double, C.c_double]
    L.lv_step_find_D.argtypes = [vp]
    L.lv_step_viscous_step.argtypes = [vp, C.c_double, C.c_int32]
    L.lv_step_bdary_friction.argtypes = [vp, C.c_double, C.c_void_p]
    L.lv_step_find_dv.argtypes = [vp, C.c_double, C.c_double]
    L.lv_step_relaxation_step.argtypes = [vp, C.c_double, C.c_int32]
    L.lv_step_lloyd.argtypes = [vp, C.c_int32]
    L.lv_step_multiphase_projection.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_int32, i32p, i32p]
    L.lv_prof_enable.argtypes = [vp, C.c_int32]
    L.lv_prof_reset.argtypes = [vp]
    L.lv_prof_get.argtypes = [vp, C.c_int32, dp, ip]
    L.lv_launch_count.argtypes = [vp]
    L.lv_launch_count.restype = C.c_int64
    L.lv_device_bytes.argtypes = [vp]
    L.lv_device_bytes.restype = C.c_int64
    for name in SYMBOLS:
        fn = getattr(L, name)
        if name not in ("lv_last_error", "lv_launch_count", "lv_device_bytes"):
            fn.restype = C.c_int32
    _lib = L
    return L


_MESSAGES = {LV_EDESTROYED: "The Voronoi Mesh has been destroyed.", LV_ENAN: "Velocity field invalidated."}


def check(status: int, handle=None) -> None:
    if status == LV_OK:
        return
    L = load_library()
    msg = L.lv_last_error(handle)
    text = msg.decode() if msg else ""
    if not text:
        text = _MESSAGES.get(status, f"lv status {status}")
    if status == LV_EINVAL and "h must be positive" in text:
        raise ValueError(text)  # ArgumentError  neighborlist.jl:19-21
    raise LvError(status, text)


def ptr(a):
    """void* of a numpy array / torch tensor / raw int (None -> NULL)."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous()
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))
