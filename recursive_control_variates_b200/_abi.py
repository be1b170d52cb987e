"""ctypes mirror of include/rcv_b200.h and loader of librcv_b200.so.

The CUDA library is the only compute path of this package: if it cannot be
loaded every compute entry point raises (there is no CPU fallback, and the
oracle under oracle/ is never imported from here).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RCV_B200_LIB selects an alternative build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("RCV_B200_LIB") or os.path.join(_HERE, "librcv_b200.so")

RCV_PHASE_ISOTROPIC, RCV_PHASE_HG = 0, 1
RCV_ADJOINT_VOLPATHSIMPLE, RCV_ADJOINT_PRBVOLPATH = 0, 1
RCV_EMITTER_CONSTANT, RCV_EMITTER_ENVMAP, RCV_EMITTER_AREA = 0, 1, 2
RCV_LOSS_L1, RCV_LOSS_L2 = 0, 1
RCV_TRACER_WAVEFRONT, RCV_TRACER_SIMPLE, RCV_TRACER_BATCH, RCV_TRACER_POOL = 0, 1, 2, 3


RCV_MAX_SENSORS = 16
RCV_RFILTER_BOX, RCV_RFILTER_GAUSSIAN = 0, 1


RCV_MAX_QUADS = 8


class Quad(C.Structure):
    """struct rcv_quad (include/rcv_b200.h)."""
    _fields_ = [("p0", C.c_float * 3), ("eu", C.c_float * 3), ("ev", C.c_float * 3),
                ("reflectance", C.c_float * 3), ("radiance", C.c_float * 3)]


class SceneDesc(C.Structure):
    """struct rcv_scene_desc (include/rcv_b200.h)."""
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("film_width", C.c_uint32),
        ("film_height", C.c_uint32),
        ("cam_origin", C.c_float * 3),
        ("cam_axes", C.c_float * 9),
        ("cam_tan_x", C.c_float),
        ("cam_tan_y", C.c_float),
        ("cam_near_clip", C.c_float),
        ("bbox_min", C.c_float * 3),
        ("bbox_max", C.c_float * 3),
        ("sigma_t_scale", C.c_float),
        ("phase_type", C.c_int32),
        ("phase_g", C.c_float),
        ("sample_emitters", C.c_int32),
        ("env_radiance", C.c_float * 3),
        ("env_dist", C.c_float),
        ("max_depth", C.c_int32),
        ("rr_depth", C.c_int32),
        ("hide_emitters", C.c_int32),
        ("use_nee", C.c_int32),
        ("n_sensors", C.c_uint32),
        ("batch_origin", (C.c_float * 3) * RCV_MAX_SENSORS),
        ("batch_axes", (C.c_float * 9) * RCV_MAX_SENSORS),
        ("rfilter_type", C.c_int32),
        ("rfilter_stddev", C.c_float),
        ("emitter_type", C.c_int32),
        ("envmap_rgb", C.c_void_p),
        ("env_width", C.c_int32),
        ("env_height", C.c_int32),
        ("env_scale", C.c_float),
        ("env_to_world", C.c_float * 9),
        ("n_quads", C.c_uint32),
        ("area_quad", C.c_int32),
        ("quads", Quad * RCV_MAX_QUADS),
    ]


class IterHostArgs(C.Structure):
    """struct rcv_iter_host_args (include/rcv_b200.h)."""
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("sigma_t_in", C.c_void_p),
        ("albedo_in", C.c_void_p),
        ("nz", C.c_int32), ("ny", C.c_int32), ("nx", C.c_int32),
        ("ref", C.c_void_p),
        ("seed", C.c_uint32), ("spp", C.c_uint32),
        ("seed_grad", C.c_uint32), ("spp_grad", C.c_uint32),
        ("loss_kind", C.c_int32),
        ("beta1", C.c_float), ("beta2", C.c_float),
        ("warmup", C.c_int32), ("reset", C.c_int32), ("adjoint_kind", C.c_int32),
        ("img_out", C.c_void_p),
        ("loss_out", C.c_void_p),
        ("grad_sigma_t_out", C.c_void_p),
        ("grad_albedo_out", C.c_void_p),
    ]


class RcvError(RuntimeError):
    pass


_P = C.c_void_p
_u32, _i32, _u64, _f = C.c_uint32, C.c_int32, C.c_uint64, C.c_float

# name -> argtypes; every function returns int (status) unless noted
_SIGNATURES = {
    "rcv_abi_version": [],
    "rcv_device_count": [],
    "rcv_ctx_create": [C.c_int, _P, C.POINTER(_P)],
    "rcv_ctx_destroy": [_P],
    "rcv_ctx_set_stream": [_P, _P],
    "rcv_ctx_set_tracer": [_P, C.c_int],
    "rcv_adjoint_set_drt": [_P, C.c_int, C.c_int, C.c_int],
    "rcv_scene_set": [_P, C.POINTER(SceneDesc)],
    "rcv_medium_rotate": [_P, _P, _P, C.c_int, C.c_int, C.c_int],
    "rcv_medium_majorants_host": [_P, C.POINTER(_f * 2)],
    "rcv_medium_read": [_P, C.c_int, _P, _P],
    "rcv_medium_set_majorant_factor": [_P, C.c_int],
    "rcv_medium_majorant_grid_host": [_P, _P, _P, C.POINTER(C.c_int * 3)],
    "rcv_render_twostates": [_P, _u32, _u32, _u32, _u32, _P, _P],
    "rcv_render_primal": [_P, C.c_int, _u32, _u32, _u32, _u32, _P],
    "rcv_develop": [_P, _P, _P],
    "rcv_cv_update": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _u64, _i32, _i32, _i32, _f, _f, _i32],
    "rcv_loss_grad": [_P, C.c_int, _P, _P, _u64, _P, _P],
    "rcv_render_backward": [_P, C.c_int, _u32, _u32, _u32, _u32, _P],
    "rcv_render_backward_weighted": [_P, C.c_int, _u32, _u32, _u32, _u32, _P, _P],
    "rcv_film_weights": [_P, _u32, _u32, _u32, _u32, _P],
    "rcv_render_adjoint_primal": [_P, C.c_int, _u32, _u32, _u32, _u32, _P],
    "rcv_grad_zero": [_P],
    "rcv_grad_buffer": [_P, C.POINTER(_P), C.POINTER(_u64)],
    "rcv_grad_read": [_P, _P, _P],
    "rcv_adam_step": [_P, _P, _P, _P, _P, _P, _P, _i32, _f, _f, _f, _f, _f, _f, _f, _f, _f],
    "rcv_adam_flat": [_P, _P, _P, _P, _P, _u64, _i32, _f, _f, _f, _f],
    "rcv_clamp": [_P, _P, _u64, _f, _f],
    "rcv_upsample2x": [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int],
    "rcv_counters_host": [_P, C.POINTER(_u64 * 8), C.c_int],
    "rcv_cv_iteration_host": [_P, C.POINTER(IterHostArgs)],
}

EXPORTED_SYMBOLS = ["rcv_last_error"] + list(_SIGNATURES)

_lib = None


def load_library() -> C.CDLL:
    """Load librcv_b200.so (built in-tree by __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RcvError(
            f"{LIB_PATH} not found: build the CUDA library first "
            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.rcv_last_error.restype = C.c_char_p
    lib.rcv_last_error.argtypes = []
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = argtypes
    if lib.rcv_abi_version() != 4:
        raise RcvError("librcv_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load_library().rcv_last_error()
        raise RcvError(f"librcv_b200 error {status}: {msg.decode() if msg else '?'}")
