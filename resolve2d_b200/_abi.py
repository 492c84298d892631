"""ctypes view of include/r2d_abi.h: structures, status codes and the loader of libr2d_b200.so.

The library is the product: hand-written sm_100a CUDA kernels behind a C ABI.  There is no CPU fallback — if the
shared object is missing or no CUDA device is usable, loading / r2d_create fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

R2D_OK = 0
ERRORS = {
    -1: "OutOfMemory",
    -2: "InvalidRigidBodyId",
    -3: "NoSuchIdExists",
    -4: "InvalidArgument",
    -5: "NoDevice",
    -6: "CudaError",
    -7: "ColorOverflow",
    -8: "BadState",
    -9: "GridRange",
}

SHAPE_DISC, SHAPE_RECT = 0, 1
JOINT_DISTANCE, JOINT_OFFSET_DISTANCE, JOINT_FIXED_POSITION, JOINT_MOTOR = 0, 1, 2, 3
MODE_PARITY, MODE_FAST, MODE_REFERENCE_ORDER = 0, 1, 2
OPT_WARM_START, OPT_SLEEPING, OPT_SLEEP_CALLS = 1, 2, 3
KCLASS_NAMES = ["broadphase", "narrowphase", "coloring", "integrate", "solve_contacts", "solve_joints"]


class BodyOpts(C.Structure):
    _fields_ = [
        ("pos_x", C.c_float), ("pos_y", C.c_float),
        ("vel_x", C.c_float), ("vel_y", C.c_float),
        ("angle", C.c_float), ("omega", C.c_float), ("mu", C.c_float),
        ("mass_value", C.c_float), ("mass_is_density", C.c_int32),
    ]


class BodyDesc(C.Structure):
    _fields_ = [
        ("opts", BodyOpts),
        ("shape", C.c_int32), ("a", C.c_float), ("b", C.c_float), ("is_static", C.c_int32),
    ]


class JointParams(C.Structure):
    _fields_ = [("power_max", C.c_float), ("power_min", C.c_float), ("beta", C.c_float)]


class BodyState(C.Structure):
    _fields_ = [
        ("id", C.c_uint32), ("shape", C.c_int32), ("is_static", C.c_int32),
        ("pos_x", C.c_float), ("pos_y", C.c_float), ("angle", C.c_float),
        ("momentum_x", C.c_float), ("momentum_y", C.c_float), ("ang_momentum", C.c_float),
        ("force_x", C.c_float), ("force_y", C.c_float), ("torque", C.c_float),
        ("mass", C.c_float), ("inertia", C.c_float), ("mu", C.c_float),
        ("aabb_x", C.c_float), ("aabb_y", C.c_float), ("aabb_half_w", C.c_float), ("aabb_half_h", C.c_float),
        ("shape_a", C.c_float), ("shape_b", C.c_float),
    ]


class Manifold(C.Structure):
    _fields_ = [
        ("ref_id", C.c_uint32), ("inc_id", C.c_uint32), ("normal_id", C.c_uint32), ("n_points", C.c_uint32),
        ("normal_x", C.c_float), ("normal_y", C.c_float),
        ("pos_x", C.c_float * 2), ("pos_y", C.c_float * 2), ("depth", C.c_float * 2),
        ("ref_rx", C.c_float * 2), ("ref_ry", C.c_float * 2), ("inc_rx", C.c_float * 2), ("inc_ry", C.c_float * 2),
        ("color", C.c_uint32),
    ]


class StepStats(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "n_bodies", "n_buckets", "n_entries", "n_pairs", "n_manifolds", "n_points", "n_colors", "n_color_rounds",
        "n_joints", "n_joint_colors", "n_launches", "n_dropped")]


# numpy dtype twins of the structures above (bulk creation / bulk readback)
def body_desc_dtype():
    import numpy as np
    return np.dtype([
        ("pos_x", "<f4"), ("pos_y", "<f4"), ("vel_x", "<f4"), ("vel_y", "<f4"), ("angle", "<f4"), ("omega", "<f4"),
        ("mu", "<f4"), ("mass_value", "<f4"), ("mass_is_density", "<i4"),
        ("shape", "<i4"), ("a", "<f4"), ("b", "<f4"), ("is_static", "<i4"),
    ])


def manifold_dtype():
    import numpy as np
    return np.dtype([
        ("ref_id", "<u4"), ("inc_id", "<u4"), ("normal_id", "<u4"), ("n_points", "<u4"),
        ("normal_x", "<f4"), ("normal_y", "<f4"),
        ("pos_x", "<f4", 2), ("pos_y", "<f4", 2), ("depth", "<f4", 2),
        ("ref_rx", "<f4", 2), ("ref_ry", "<f4", 2), ("inc_rx", "<f4", 2), ("inc_ry", "<f4", 2),
        ("color", "<u4"),
    ])


class R2DError(RuntimeError):
    def __init__(self, code: int, what: str, detail: str = ""):
        self.code = code
        self.name = ERRORS.get(code, f"status {code}")
        super().__init__(f"{what}: {self.name}" + (f" ({detail})" if detail else ""))


_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libr2d_b200.so")
_lib = None

# symbol -> (restype, argtypes); every symbol include/r2d_abi.h declares
_P = C.c_void_p
_F = C.c_float
_U32 = C.c_uint32
_SZ = C.c_size_t
_FP = C.POINTER(C.c_float)
_UP = C.POINTER(C.c_uint32)
SIGNATURES = {
    "r2d_abi_version": (C.c_int, []),
    "r2d_last_error": (C.c_char_p, []),
    "r2d_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "r2d_create": (C.c_int, [_F, _U32, C.c_int, C.POINTER(_P)]),
    "r2d_destroy": (C.c_int, [_P]),
    "r2d_clear": (C.c_int, [_P]),
    "r2d_set_mode": (C.c_int, [_P, C.c_int]),
    "r2d_set_stream": (C.c_int, [_P, _P]),
    "r2d_set_option": (C.c_int, [_P, C.c_int, _U32]),
    "r2d_batch_set_option": (C.c_int, [_P, C.c_int, _U32]),
    "r2d_set_reorder_interval": (C.c_int, [_P, _U32]),
    "r2d_reorder": (C.c_int, [_P]),
    "r2d_make_disc": (C.c_int, [_P, C.POINTER(BodyOpts), _F, _UP]),
    "r2d_make_rect": (C.c_int, [_P, C.POINTER(BodyOpts), _F, _F, _UP]),
    "r2d_make_bodies": (C.c_int, [_P, _P, _SZ, _UP]),
    "r2d_make_gravity": (C.c_int, [_P, _F]),
    "r2d_make_distance_joint": (C.c_int, [_P, C.POINTER(JointParams), _U32, _U32, _F, C.POINTER(_SZ)]),
    "r2d_make_offset_distance_joint": (C.c_int, [_P, C.POINTER(JointParams), _U32, _U32, _F, _F, _F, _F, _F, C.POINTER(_SZ)]),
    "r2d_make_fixed_position_joint": (C.c_int, [_P, C.POINTER(JointParams), _U32, _F, _F, C.POINTER(_SZ)]),
    "r2d_make_motor_joint": (C.c_int, [_P, C.POINTER(JointParams), _U32, _F, C.POINTER(_SZ)]),
    "r2d_exclude_pair": (C.c_int, [_P, _U32, _U32]),
    "r2d_remove_body": (C.c_int, [_P, _U32]),
    "r2d_process": (C.c_int, [_P, _F, _U32, _U32]),
    "r2d_step": (C.c_int, [_P, _F, _U32, _U32]),
    "r2d_process_read": (C.c_int, [_P, _F, _U32, _U32, _P, _P, _P, _P, _P, _P, _SZ]),
    "r2d_synchronize": (C.c_int, [_P]),
    "r2d_num_bodies": (C.c_int, [_P, C.POINTER(_SZ)]),
    "r2d_body_id_at": (C.c_int, [_P, _SZ, _UP]),
    "r2d_body_get": (C.c_int, [_P, _U32, C.POINTER(BodyState)]),
    "r2d_body_set_static": (C.c_int, [_P, _U32, C.c_int]),
    "r2d_body_set_pos": (C.c_int, [_P, _U32, _F, _F]),
    "r2d_body_set_angle": (C.c_int, [_P, _U32, _F]),
    "r2d_body_set_momentum": (C.c_int, [_P, _U32, _F, _F]),
    "r2d_body_set_ang_momentum": (C.c_int, [_P, _U32, _F]),
    "r2d_body_set_force": (C.c_int, [_P, _U32, _F, _F]),
    "r2d_body_set_torque": (C.c_int, [_P, _U32, _F]),
    "r2d_read_bodies": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _SZ]),
    "r2d_write_forces": (C.c_int, [_P, _P, _SZ]),
    "r2d_read_pairs": (C.c_int, [_P, _P, _P, _SZ, C.POINTER(_SZ)]),
    "r2d_read_manifolds": (C.c_int, [_P, _P, _SZ, C.POINTER(_SZ)]),
    "r2d_read_joint_order": (C.c_int, [_P, _P, _P, _SZ, C.POINTER(_SZ)]),
    "r2d_get_stats": (C.c_int, [_P, C.POINTER(StepStats)]),
    "r2d_batch_create": (C.c_int, [_U32, _F, _U32, C.c_int, C.POINTER(_P)]),
    "r2d_batch_destroy": (C.c_int, [_P]),
    "r2d_batch_world": (C.c_int, [_P, _U32, C.POINTER(_P)]),
    "r2d_batch_num_worlds": (C.c_int, [_P, _UP]),
    "r2d_batch_set_mode": (C.c_int, [_P, C.c_int]),
    "r2d_batch_set_stream": (C.c_int, [_P, _P]),
    "r2d_batch_set_reorder_interval": (C.c_int, [_P, _U32]),
    "r2d_batch_reorder": (C.c_int, [_P]),
    "r2d_batch_process": (C.c_int, [_P, _F, _U32, _U32]),
    "r2d_batch_process_read": (C.c_int, [_P, _F, _U32, _U32, _P, _P, _P, _P, _P, _P, _SZ]),
    "r2d_batch_synchronize": (C.c_int, [_P]),
    "r2d_batch_num_bodies": (C.c_int, [_P, C.POINTER(_SZ)]),
    "r2d_batch_read_bodies": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _SZ]),
    "r2d_batch_write_forces": (C.c_int, [_P, _P, _SZ]),
    "r2d_batch_get_stats": (C.c_int, [_P, C.POINTER(StepStats)]),
    "r2d_sharded_create": (C.c_int, [_U32, C.POINTER(C.c_int), _U32, _F, _U32, C.POINTER(_P)]),
    "r2d_sharded_destroy": (C.c_int, [_P]),
    "r2d_sharded_num_worlds": (C.c_int, [_P, _UP]),
    "r2d_sharded_num_shards": (C.c_int, [_P, _UP]),
    "r2d_sharded_shard": (C.c_int, [_P, _U32, C.POINTER(_P), _UP, _UP]),
    "r2d_sharded_world": (C.c_int, [_P, _U32, C.POINTER(_P)]),
    "r2d_sharded_set_mode": (C.c_int, [_P, C.c_int]),
    "r2d_sharded_reorder": (C.c_int, [_P]),
    "r2d_sharded_process": (C.c_int, [_P, _F, _U32, _U32]),
    "r2d_sharded_process_read": (C.c_int, [_P, _F, _U32, _U32, _P, _P, _P, _P, _P, _P, _SZ]),
    "r2d_sharded_synchronize": (C.c_int, [_P]),
    "r2d_sharded_num_bodies": (C.c_int, [_P, C.POINTER(_SZ)]),
    "r2d_sharded_read_bodies": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _SZ]),
    "r2d_sharded_write_forces": (C.c_int, [_P, _P, _SZ]),
    "r2d_sharded_get_stats": (C.c_int, [_P, C.POINTER(StepStats)]),
    "r2d_batch_profile_enable": (C.c_int, [_P, C.c_int]),
    "r2d_profile_enable": (C.c_int, [_P, C.c_int]),
    "r2d_batch_profile_read": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]),
    "r2d_profile_read": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]),
}


def bind(lib, signatures=SIGNATURES, prefix_from="r2d_", prefix_to="r2d_"):
    """Attach restype/argtypes; raises AttributeError naming the first missing symbol."""
    for name, (res, args) in signatures.items():
        sym = prefix_to + name[len(prefix_from):] if name.startswith(prefix_from) else name
        fn = getattr(lib, sym)
        fn.restype = res
        fn.argtypes = args
    return lib


def load_library(path: str | None = None):
    """Load libr2d_b200.so (built by `make -C resolve2d_b200/csrc` or __graft_entry__.build()).  No fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("R2D_B200_LIB", LIB_PATH)
    if not os.path.exists(p):
        raise ImportError(
            f"{p} not found: the CUDA extension is the only implementation of this package (there is no CPU "
            "fallback). Build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C resolve2d_b200/csrc`.")
    lib = bind(C.CDLL(p))
    if path is None:
        _lib = lib
    return lib


def check(lib, status: int, what: str):
    if status != R2D_OK:
        detail = ""
        try:
            detail = (lib.r2d_last_error() or b"").decode("utf-8", "replace")
        except Exception:  # pragma: no cover - oracle library has no r2d_last_error
            pass
        raise R2DError(status, what, detail)
