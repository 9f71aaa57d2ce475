"""ctypes binding of ``libartist_b200.so`` (the C ABI in ``include/artist_b200.h``).

The product path has no CPU or eager fallback: if the library is missing or a call fails,
an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

from ._build import LIB_OVERRIDE, LIB_PATH as _DEFAULT_LIB_PATH

LIB_PATH = LIB_OVERRIDE or _DEFAULT_LIB_PATH

ABI_VERSION = 4
TRIG_SINCOSF, TRIG_TABLE, TRIG_POLY = 0, 1, 2
FLAG_FP32_ACCUM = 1
FLAG_ONE_CTA_PER_SAMPLE = 2

c_float_p = C.c_void_p  # raw device/host addresses (tensor.data_ptr())
c_int_p = C.c_void_p


class Targets(C.Structure):
    _fields_ = [
        ("n_planar", C.c_int32), ("n_cyl", C.c_int32),
        ("planar_centers", c_float_p), ("planar_normals", c_float_p), ("planar_dims", c_float_p),
        ("cyl_centers", c_float_p), ("cyl_normals", c_float_p), ("cyl_axes", c_float_p),
        ("cyl_radii", c_float_p), ("cyl_heights", c_float_p), ("cyl_opening", c_float_p),
    ]


class Blockers(C.Structure):
    _fields_ = [
        ("n_blockers", C.c_int32), ("max_candidates", C.c_int32), ("prims", c_float_p), ("cand_idx", c_int_p),
        ("cand_count", c_int_p), ("softness", C.c_float), ("alpha", C.c_float), ("ray_origin_offset", C.c_float),
        ("epsilon", C.c_float), ("cull_angle", C.c_float),
    ]


class TraceArgs(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n_samples", C.c_int32), ("n_points", C.c_int32), ("n_rays", C.c_int32),
        ("res_e", C.c_int32), ("res_u", C.c_int32), ("n_local", C.c_int32), ("local_rows", c_int_p),
        ("points", c_float_p), ("normals", c_float_p), ("incident", c_float_p), ("distortions", c_float_p),
        ("trig", c_float_p), ("target_idx", c_int_p), ("targets", Targets), ("blockers", Blockers),
        ("ray_magnitude", C.c_float), ("one_minus_extinction", C.c_float), ("reflectivity", C.c_float),
        ("scatter_sigma", C.c_float), ("trig_mode", C.c_int32), ("flags", C.c_int32),
        ("flux", c_float_p), ("intercept", c_float_p), ("on_target", c_float_p), ("blocking", c_float_p),
        ("dbg_be", c_float_p), ("dbg_bu", c_float_p), ("dbg_t", c_float_p), ("dbg_lambert", c_float_p),
        ("stats", C.c_void_p), ("orientations", c_float_p), ("windows", c_int_p), ("distortions_planar", c_float_p),
        ("src_rows", c_int_p),
    ]


class TraceBwdArgs(C.Structure):
    _fields_ = [("fwd", TraceArgs), ("grad_flux", c_float_p), ("grad_flux_stride", C.c_int64),
                ("grad_points", c_float_p), ("grad_normals", c_float_p), ("grad_prims", c_float_p),
                ("grad_orientations", c_float_p), ("grad_prims_scratch", c_float_p), ("grad_prims_scratch_floats", C.c_int64)]


class NurbsArgs(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n_surfaces", C.c_int32), ("n_facets", C.c_int32), ("n_eval", C.c_int32),
        ("n_ctrl_u", C.c_int32), ("n_ctrl_v", C.c_int32), ("degree_u", C.c_int32), ("degree_v", C.c_int32),
        ("control_points", c_float_p), ("eval_points", c_float_p), ("eval_stride_n", C.c_int64),
        ("eval_stride_f", C.c_int64), ("knots_u", c_float_p), ("knots_v", c_float_p), ("canting", c_float_p),
        ("facet_translations", c_float_p), ("grid_u", C.c_int32), ("grid_v", C.c_int32), ("points", c_float_p),
        ("normals", c_float_p),
    ]


class NurbsBwdArgs(C.Structure):
    _fields_ = [("fwd", NurbsArgs), ("grad_points", c_float_p), ("grad_normals", c_float_p),
                ("grad_control_points", c_float_p)]


class KinematicsArgs(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n", C.c_int32), ("linear_actuators", C.c_int32),
        ("positions", c_float_p), ("translation_dev", c_float_p), ("rotation_dev", c_float_p),
        ("actuator_non_opt", c_float_p), ("actuator_opt", c_float_p), ("orientation_offset", c_float_p),
    ]


class HostTraceArgs(C.Structure):
    _fields_ = [
        ("dev", TraceArgs), ("h_points", c_float_p), ("h_normals", c_float_p), ("h_incident", c_float_p),
        ("h_target_idx", c_int_p), ("d_target_bitmaps", c_float_p), ("h_target_bitmaps", c_float_p),
        ("h_factors", c_float_p),
    ]


EXPORTS = {
    "ab200_trace_fwd": ([C.POINTER(TraceArgs), C.c_void_p], C.c_int32),
    "ab200_trace_bwd": ([C.POINTER(TraceBwdArgs), C.c_void_p], C.c_int32),
    "ab200_bitmaps_per_target": ([c_float_p, c_int_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_float_p, C.c_void_p], C.c_int32),
    "ab200_nurbs_fwd": ([C.POINTER(NurbsArgs), C.c_void_p], C.c_int32),
    "ab200_nurbs_bwd": ([C.POINTER(NurbsBwdArgs), C.c_void_p], C.c_int32),
    "ab200_kinematics_fwd": ([C.POINTER(KinematicsArgs), c_float_p, c_float_p, C.c_void_p], C.c_int32),
    "ab200_kinematics_bwd": ([C.POINTER(KinematicsArgs), c_float_p, c_float_p, c_float_p, c_float_p, c_float_p,
                              c_float_p, c_float_p, C.c_void_p], C.c_int32),
    "ab200_kinematics_align_incident": ([C.POINTER(KinematicsArgs), c_float_p, c_float_p, C.c_int32, C.c_float,
                                         c_float_p, c_float_p, c_float_p, C.c_void_p], C.c_int32),
    "ab200_align_fwd": ([c_float_p, c_float_p, c_float_p, c_int_p, C.c_int32, C.c_int32, c_float_p, c_float_p,
                         C.c_void_p], C.c_int32),
    "ab200_align_bwd": ([c_float_p, c_float_p, c_float_p, c_int_p, C.c_int32, C.c_int32, c_float_p, c_float_p,
                         c_float_p, c_float_p, c_float_p, C.c_void_p], C.c_int32),
    "ab200_trace_host": ([C.POINTER(HostTraceArgs), C.c_void_p], C.c_int32),
    "ab200_flux_moments": ([c_float_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_float_p, C.c_void_p], C.c_int32),
    "ab200_flux_moments_bwd": ([c_float_p, c_float_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_float_p, C.c_void_p], C.c_int32),
    "ab200_flux_crop_fwd": ([c_float_p, c_float_p, C.c_int32, C.c_int32, C.c_int32, c_float_p, c_float_p, C.c_void_p], C.c_int32),
    "ab200_flux_crop_bwd": ([c_float_p, c_float_p, c_float_p, c_float_p, C.c_int32, C.c_int32, C.c_int32, c_float_p, c_float_p,
                             C.c_void_p], C.c_int32),
    "ab200_flux_loss_fwd": ([c_float_p, c_float_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_float_p, c_float_p, C.c_void_p],
                            C.c_int32),
    "ab200_flux_loss_bwd": ([c_float_p, c_float_p, c_float_p, c_float_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_float_p,
                            C.c_void_p], C.c_int32),
    "ab200_reflect": ([c_float_p, c_float_p, C.c_int32, C.c_int32, c_float_p, C.c_void_p], C.c_int32),
    "ab200_scatter_rays": ([c_float_p, c_float_p, c_float_p, C.c_int32, C.c_int32, C.c_int32, c_float_p, C.c_void_p], C.c_int32),
    "ab200_line_intersections": ([c_float_p, c_float_p, c_float_p, C.POINTER(Targets), c_int_p, C.c_int32, C.c_int32,
                                  C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_float_p, c_float_p, c_float_p, c_float_p,
                                  C.c_void_p], C.c_int32),
    "ab200_bilinear_splatting": ([c_float_p, c_float_p, c_float_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, c_float_p,
                                  C.c_void_p], C.c_int32),
    "ab200_blocking_pack": ([c_float_p, c_float_p, c_float_p, C.c_int32, C.c_float, c_float_p, C.c_void_p], C.c_int32),
    "ab200_blocking_candidates": ([c_float_p, C.c_int32, c_int_p, c_float_p, c_float_p, C.c_int32, C.c_float, C.c_int32,
                                   c_int_p, c_int_p, c_int_p, C.c_void_p], C.c_int32),
    "ab200_sample_distortions": ([c_float_p, C.c_int64, C.c_uint64, C.c_uint64, C.c_float, C.c_float, C.c_float, C.c_float,
                                  C.c_int32, C.c_int32, C.POINTER(C.c_uint64), c_float_p, C.c_void_p], C.c_int32),
    "ab200_deinterleave_distortions": ([c_float_p, C.c_int64, c_float_p, C.c_void_p], C.c_int32),
    "ab200_replica_sum": ([c_float_p, c_int_p, C.c_int32, C.c_int64, c_float_p, C.c_void_p], C.c_int32),
    "ab200_abi_version": ([], C.c_int32),
    "ab200_kernel_launch_count": ([], C.c_int64),
    "ab200_error_string": ([C.c_int32], C.c_char_p),
    "ab200_last_error_detail": ([], C.c_char_p),
    "ab200_debug_const_div": ([c_float_p, C.c_int32, C.c_float, c_float_p, c_float_p, C.c_void_p], C.c_int32),
    "ab200_debug_div_regular": ([c_float_p, c_float_p, C.c_int32, c_float_p, c_float_p, C.c_void_p], C.c_int32),
    "ab200_debug_trig": ([c_float_p, C.c_int32, C.c_int32, c_float_p, c_float_p, C.c_void_p], C.c_int32),
}

_lib = None
launch_count = 0  # number of C-ABI compute calls issued through this binding (bench.py reports it)


class Ab200Error(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load the library (once).  Raises if it has not been built - there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Ab200Error(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). artist_b200 has no CPU/eager fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (argtypes, restype) in EXPORTS.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = restype
        if handle.ab200_abi_version() != ABI_VERSION:
            raise Ab200Error("libartist_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        h = lib()
        raise Ab200Error(f"{what} failed: {h.ab200_error_string(rc).decode()} - {h.ab200_last_error_detail().decode()}")


# Optional per-entry-point device timing (bench.py): CUDA events on the launching stream.
timing_enabled = False
timing_events: dict[str, list] = {}


def call(name: str, *args) -> None:
    global launch_count
    launch_count += 1
    if timing_enabled:
        import torch

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(getattr(lib(), name)(*args), name)
        e1.record()
        timing_events.setdefault(name, []).append((e0, e1))
    else:
        check(getattr(lib(), name)(*args), name)


def kernel_launches() -> int:
    return int(lib().ab200_kernel_launch_count())
