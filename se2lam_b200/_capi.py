"""ctypes binding of the C ABI declared in include/se2gpu.h (the drop-in boundary).

Loading fails loudly if the CUDA library has not been built: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

KP_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"), ("response", "f4"),
                     ("octave", "i4"), ("class_id", "i4")])
assert KP_DTYPE.itemsize == 28
BA_STATS_DTYPE = np.dtype([("chi2_before", "f8"), ("chi2_after", "f8"), ("lambda", "f8"), ("rho", "f8"),
                           ("trials", "i4"), ("accepted", "i4"), ("terminate", "i4"), ("pad", "i4")])

ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p)


class GridParams(C.Structure):
    _fields_ = [("min_x", C.c_float), ("min_y", C.c_float), ("inv_w", C.c_float), ("inv_h", C.c_float)]


class BowKF(C.Structure):
    _fields_ = [("angle", C.c_void_p), ("desc", C.c_void_p), ("has_mp", C.c_void_p), ("n", C.c_int),
                ("node", C.c_void_p), ("n_node", C.c_int), ("ptr", C.c_void_p), ("feat", C.c_void_p)]


class Se2GpuError(RuntimeError):
    pass


_lib = None

# every symbol include/se2gpu.h declares (tests/test_abi.py checks the header against this list)
PEER_HANDLE_BYTES = 128   # SE2GPU_BA_PEER_HANDLE_BYTES

SYMBOLS = [
    "se2gpu_device_count", "se2gpu_last_error", "se2gpu_launch_count",
    "se2gpu_orb_create", "se2gpu_orb_destroy", "se2gpu_orb_extract", "se2gpu_orb_extract_device", "se2gpu_orb_submit", "se2gpu_orb_wait",
    "se2gpu_orb_level_dims", "se2gpu_orb_get_level", "se2gpu_orb_profile", "se2gpu_orb_profile_read",
    "se2gpu_orb_debug_nth_element", "se2gpu_orb_set_undistort", "se2gpu_orb_debug_undistort_map",
    "se2gpu_hamming_distance", "se2gpu_match_by_window", "se2gpu_match_by_projection", "se2gpu_search_by_bow",
    "se2gpu_matcher_create", "se2gpu_matcher_destroy", "se2gpu_match_by_window_device", "se2gpu_keypoints_to_points_device",
    "se2gpu_match_by_projection_device", "se2gpu_matcher_match_by_window", "se2gpu_matcher_match_by_projection",
    "se2gpu_matcher_search_by_bow", "se2gpu_matcher_profile", "se2gpu_matcher_profile_read", "se2gpu_matcher_last_rounds",
    "se2gpu_ba_create", "se2gpu_ba_destroy", "se2gpu_ba_set_problem", "se2gpu_ba_optimize", "se2gpu_ba_get",
    "se2gpu_ba_set_shard", "se2gpu_ba_peer_export", "se2gpu_ba_peer_import", "se2gpu_ba_set_stream", "se2gpu_ba_debug_system", "se2gpu_ba_reset", "se2gpu_ba_profile",
    "se2gpu_ba_profile_read", "se2gpu_ba_set_mode", "se2gpu_ba_get_f32", "se2gpu_ba_build_information", "se2gpu_ba_optimize_from", "se2gpu_ba_peer_attach_local",
    "se2gpu_voc_create", "se2gpu_voc_destroy", "se2gpu_voc_transform", "se2gpu_voc_transform_device", "se2gpu_median_descriptor",
]


def lib_path() -> str:
    return _build.LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise Se2GpuError(
            f"{path} is missing: build it with `python -m se2lam_b200.build` (or __graft_entry__.build()). "
            "se2lam_b200 has no CPU fallback.")
    L = C.CDLL(path)
    vp, i, f, d, sz = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_size_t
    L.se2gpu_device_count.restype = i
    L.se2gpu_last_error.restype = C.c_char_p
    L.se2gpu_launch_count.restype = C.c_ulonglong
    L.se2gpu_orb_create.restype = vp
    L.se2gpu_orb_create.argtypes = [i, f, i, i, i, i, i, i]
    L.se2gpu_orb_destroy.argtypes = [vp]
    L.se2gpu_orb_extract.argtypes = [vp, vp, i, i, i, i, sz, vp, vp, vp]
    L.se2gpu_orb_submit.argtypes = [vp, vp, i, i, i, i, sz, vp, vp, vp]
    L.se2gpu_orb_wait.argtypes = [vp]
    L.se2gpu_orb_extract_device.argtypes = [vp, vp, i, i, i, i, sz, vp, vp, vp, vp]
    L.se2gpu_orb_level_dims.argtypes = [vp, i, C.POINTER(i), C.POINTER(i), C.POINTER(i)]
    L.se2gpu_orb_get_level.argtypes = [vp, i, i, i, vp]
    L.se2gpu_orb_profile.argtypes = [vp, i]
    L.se2gpu_orb_profile_read.argtypes = [vp, vp, vp]
    L.se2gpu_orb_debug_nth_element.argtypes = [vp, vp, vp, i, i]
    L.se2gpu_orb_set_undistort.argtypes = [vp, vp, vp, i]
    L.se2gpu_orb_debug_undistort_map.argtypes = [vp, vp, i, i, i, vp, vp]
    L.se2gpu_ba_reset.argtypes = [vp]
    L.se2gpu_ba_peer_export.argtypes = [vp, vp]
    L.se2gpu_ba_peer_import.argtypes = [vp, vp, i]
    L.se2gpu_ba_profile.argtypes = [vp, i]
    L.se2gpu_ba_set_mode.argtypes = [vp, i]
    L.se2gpu_ba_profile_read.argtypes = [vp, vp, vp]
    L.se2gpu_hamming_distance.argtypes = [vp, vp, i, vp, i]
    L.se2gpu_match_by_window.argtypes = [vp, vp, i, vp, vp, i, vp, GridParams, i, i, i, i, f, vp, i]
    L.se2gpu_match_by_projection.argtypes = [vp, vp, i, vp, vp, vp, i, vp, vp, GridParams, i, i, f, vp, i]
    L.se2gpu_search_by_bow.argtypes = [C.POINTER(BowKF), C.POINTER(BowKF), i, f, i, vp, i]
    L.se2gpu_matcher_create.restype = vp
    L.se2gpu_matcher_create.argtypes = [i, i, i]
    L.se2gpu_matcher_destroy.argtypes = [vp]
    L.se2gpu_match_by_window_device.argtypes = [vp, vp, vp, i, vp, vp, vp, i, vp, vp, GridParams, i, i, i, i, f, vp, vp, vp]
    L.se2gpu_keypoints_to_points_device.argtypes = [vp, i, vp, vp, vp]
    L.se2gpu_match_by_projection_device.argtypes = [vp, vp, vp, i, vp, vp, vp, vp, i, vp, vp, GridParams, i, i, f, vp, vp, vp]
    L.se2gpu_matcher_match_by_window.argtypes = [vp, vp, vp, i, vp, vp, i, vp, GridParams, i, i, i, i, f, vp]
    L.se2gpu_matcher_match_by_projection.argtypes = [vp, vp, vp, i, vp, vp, vp, i, vp, vp, GridParams, i, i, f, vp]
    L.se2gpu_matcher_search_by_bow.argtypes = [vp, C.POINTER(BowKF), C.POINTER(BowKF), i, f, i, vp]
    L.se2gpu_matcher_profile.argtypes = [vp, i]
    L.se2gpu_matcher_profile_read.argtypes = [vp, vp, vp]
    L.se2gpu_matcher_last_rounds.argtypes = [vp, C.POINTER(i), C.POINTER(i)]
    L.se2gpu_ba_create.restype = vp
    L.se2gpu_ba_create.argtypes = [i, i, i, i, i]
    L.se2gpu_ba_destroy.argtypes = [vp]
    L.se2gpu_ba_set_problem.argtypes = [vp, i, i, i, i] + [vp] * 11 + [d, d, d, vp, d]
    L.se2gpu_ba_optimize.argtypes = [vp, i, vp, vp, vp, vp]
    L.se2gpu_ba_get.argtypes = [vp, vp, vp]
    L.se2gpu_ba_set_shard.argtypes = [vp, i, i, ALLREDUCE_FN, vp]
    L.se2gpu_ba_set_stream.argtypes = [vp, vp]
    L.se2gpu_ba_debug_system.argtypes = [vp, d] + [vp] * 10
    L.se2gpu_ba_get_f32.argtypes = [vp, vp, vp]
    L.se2gpu_ba_optimize_from.argtypes = [vp, i, i, vp, vp, vp, vp]
    L.se2gpu_ba_peer_attach_local.argtypes = [vp, i]
    L.se2gpu_ba_build_information.argtypes = [i, i, i, vp, vp, vp, vp, vp, vp, vp, vp, i, f, f, f, vp, i]
    L.se2gpu_voc_create.restype = vp
    L.se2gpu_voc_create.argtypes = [i, vp, vp, vp, vp, vp, i, i]
    L.se2gpu_voc_destroy.argtypes = [vp]
    L.se2gpu_voc_transform.argtypes = [vp, vp, i, i, vp, vp, vp]
    L.se2gpu_voc_transform_device.argtypes = [vp, vp, i, i, vp, vp, vp, vp]
    L.se2gpu_median_descriptor.argtypes = [vp, vp, i, vp, vp, i]
    _lib = L
    return L


def last_error() -> str:
    return lib().se2gpu_last_error().decode()


def check(rc: int, what: str) -> int:
    if rc < 0:
        raise Se2GpuError(f"{what} failed ({rc}): {last_error()}")
    return rc


def ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):  # torch tensor
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(int(a))
