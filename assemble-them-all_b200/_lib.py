"""ctypes loader of libsdfcontact_b200.so (the C-ABI in include/sdfcontact_b200.h).

There is no fallback: if the CUDA library is missing this raises, and every
compute call raises when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SDFC_LIB_PATH") or os.path.join(HERE, "libsdfcontact_b200.so")  # override: tuning builds

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)
c_void_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); mirrors include/sdfcontact_b200.h one to one
SIGNATURES = {
    "sdfc_version": (C.c_int, []),
    "sdfc_last_error": (C.c_char_p, []),
    "sdfc_device_count": (C.c_int, []),
    "sdfc_pinned_alloc": (C.c_int, [C.c_size_t, c_void_pp]),
    "sdfc_pinned_free": (C.c_int, [C.c_void_p]),
    "sdfc_context_create": (C.c_int, [C.c_int, c_void_pp]),
    "sdfc_context_destroy": (C.c_int, [C.c_void_p]),
    "sdfc_context_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sdfc_context_synchronize": (C.c_int, [C.c_void_p]),
    "sdfc_context_launch_count": (C.c_int64, [C.c_void_p]),
    "sdfc_context_enable_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "sdfc_context_read_timing": (C.c_int, [C.c_void_p, c_double_p, C.POINTER(C.c_int64), C.c_int]),
    "sdfc_measure_peaks": (C.c_int, [C.c_void_p, c_double_p]),
    "sdfc_selftest_exact_div": (C.c_int, [C.c_void_p, C.c_int64, C.c_uint64, C.c_double, C.POINTER(C.c_int64)]),
    "sdfc_sdf_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, C.c_double, C.c_double,
                                  C.c_double, C.c_double, c_double_p, c_void_pp]),
    "sdfc_sdf_set_col_th": (C.c_int, [C.c_void_p, C.c_double]),
    "sdfc_sdf_is_fp32": (C.c_int, [C.c_void_p]),
    "sdfc_sdf_layout_kind": (C.c_int, [C.c_void_p]),
    "sdfc_sdf_destroy": (C.c_int, [C.c_void_p]),
    "sdfc_sdf_distance": (C.c_int, [C.c_void_p, C.c_int64, c_double_p, c_double_p]),
    "sdfc_sdf_min_distance": (C.c_int, [C.c_void_p, C.c_int64, c_double_p, c_double_p]),
    "sdfc_sdf_stencil": (C.c_int, [C.c_void_p, C.c_int64, c_double_p, C.c_int, c_double_p, c_double_p, c_double_p]),
    "sdfc_sdf_distance_device": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sdfc_make_level_set3": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                       C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sdfc_sdf_layout": (C.c_int, [C.c_int, c_double_p, C.c_double, C.c_int, c_double_p, C.c_void_p, C.c_void_p, c_int32_p]),
    "sdfc_sdf_save_text": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, C.c_double, C.c_double,
                                     C.c_double, c_double_p]),
    "sdfc_sdf_load_text_header": (C.c_int, [C.c_char_p, c_int32_p]),
    "sdfc_sdf_load_text": (C.c_int, [C.c_char_p, c_double_p, c_double_p, c_double_p, c_double_p, C.c_int64]),
    "sdfc_points_create": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_int, c_void_pp]),
    "sdfc_points_count": (C.c_int, [C.c_void_p]),
    "sdfc_points_destroy": (C.c_int, [C.c_void_p]),
    "sdfc_pair_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double,
                                   c_void_pp]),
    "sdfc_pair_set_stiffness": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "sdfc_pair_set_friction": (C.c_int, [C.c_void_p, C.c_double]),
    "sdfc_pair_set_damping": (C.c_int, [C.c_void_p, C.c_double]),
    "sdfc_pair_set_scale": (C.c_int, [C.c_void_p, C.c_double]),
    "sdfc_pair_set_contact_scale": (C.c_int, [C.c_void_p, C.c_double]),
    "sdfc_pair_destroy": (C.c_int, [C.c_void_p]),
    "sdfc_pair_query": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_int32_p, c_uint8_p, c_int32_p]),
    "sdfc_pair_force": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_int32_p]),
    "sdfc_pair_force_deriv": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                        c_double_p, c_double_p, c_int32_p]),
    "sdfc_scatter_force": (None, [C.c_int, C.c_int, c_double_p, c_double_p]),
    "sdfc_scatter_deriv": (None, [C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    "sdfc_batch_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_void_pp, c_int32_p, c_int32_p, c_void_pp]),
    "sdfc_batch_destroy": (C.c_int, [C.c_void_p]),
    "sdfc_batch_eval": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int, c_double_p, c_double_p, c_double_p,
                                  c_int32_p]),
    "sdfc_batch_eval_blocks": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int, c_double_p, c_double_p, c_double_p,
                                         c_int32_p, c_double_p, c_int32_p]),
    "sdfc_batch_eval_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]),
    "sdfc_batch_detect_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "sdfc_batch_get_active": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_uint8_p]),
    "sdfc_batch_queries_per_eval": (C.c_int64, [C.c_void_p]),
    "sdfc_ground_create": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, C.c_double, C.c_double, C.c_double, C.c_double, c_void_pp]),
    "sdfc_ground_set_transform": (C.c_int, [C.c_void_p, c_double_p]),
    "sdfc_ground_set_stiffness": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "sdfc_ground_set_friction": (C.c_int, [C.c_void_p, C.c_double]),
    "sdfc_ground_set_damping": (C.c_int, [C.c_void_p, C.c_double]),
    "sdfc_ground_destroy": (C.c_int, [C.c_void_p]),
    "sdfc_ground_force": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_int32_p]),
    "sdfc_ground_force_deriv": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_int32_p]),
    "sdfc_batch_set_grounds": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), c_int32_p]),
    "sdfc_batch_get_ground_active": (C.c_int, [C.c_void_p, c_int32_p]),
    "sdfc_sim_get_ground_history": (C.c_int, [C.c_void_p, c_uint8_p]),
    "sdfc_sim_create": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, C.c_double, c_void_pp]),
    "sdfc_sim_destroy": (C.c_int, [C.c_void_p]),
    "sdfc_sim_ndof_r": (C.c_int, [C.c_void_p]),
    "sdfc_sim_set_state": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "sdfc_sim_get_state": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "sdfc_sim_set_external_force": (C.c_int, [C.c_void_p, c_double_p]),
    "sdfc_sim_clear_contact_history": (C.c_int, [C.c_void_p]),
    "sdfc_sim_set_active": (C.c_int, [C.c_void_p, c_uint8_p]),
    "sdfc_sim_forward": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p]),
    "sdfc_sim_get_status": (C.c_int, [C.c_void_p, c_uint8_p, c_uint8_p]),
    "sdfc_sim_get_poses": (C.c_int, [C.c_void_p, c_double_p]),
    "sdfc_batch_get_brick_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "sdfc_batch_get_skin_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "sdfc_sim_get_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
}


class BodyDesc(C.Structure):
    """struct sdfc_body_desc (include/sdfcontact_b200.h)."""
    _fields_ = [("joint_type", C.c_int32), ("reserved", C.c_int32), ("E_pj_0", C.c_double * 16),
                ("E_ji", C.c_double * 16), ("inertia", C.c_double * 6), ("mass", C.c_double)]

_lib = None


class SdfContactError(RuntimeError):
    """Mirrors the reference's `throw std::string` (Common.h:58-61)."""


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SdfContactError(
            f"{LIB_PATH} is missing: build it with `python assemble-them-all_b200/build.py` "
            "(there is no CPU fallback for the contact path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    # the same entry with untyped pointer arguments: callers that already hold addresses (PinnedArray.ptr) skip the
    # per-call construction of typed ctypes pointers (several microseconds each)
    proto = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)
    lib.sdfc_batch_eval_addr = proto(("sdfc_batch_eval", lib))
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load().sdfc_last_error()
        raise SdfContactError(f"[Error] {msg.decode() if msg else 'unknown'} (status {status})")
