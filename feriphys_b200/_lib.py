"""ctypes binding of ``libferiphys_cuda.so`` (the C ABI in include/feriphys_cuda.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device
is usable, calls fail loudly with :class:`FeriphysError`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libferiphys_cuda.so")

OK = 0
METHOD_AUTO, METHOD_ALLPAIRS, METHOD_GRID, METHOD_SMALL = 0, 1, 2, 3
NUMERICS_EXACT, NUMERICS_FAST = 0, 1
STATEFUL_TEST_POINT, STATEFUL_TEST_EXAMPLEFN, STATEFUL_SPRINGY_POINT, STATEFUL_RIGIDBODY, STATEFUL_BOID = 1, 2, 3, 4, 5
STATUS_STEER_NEGATIVE, STATUS_STEER_NAN_OVF = 1, 2


class FeriphysError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"feriphys-cuda error {code}: {msg}")
        self.code = code


class FpConfig(C.Structure):
    """``fp_config`` == ``flocking::Config`` (flocking.rs:15-34)."""
    _fields_ = [
        ("dt", C.c_float),
        ("avoidance_factor", C.c_float),
        ("centering_factor", C.c_float),
        ("velocity_matching_factor", C.c_float),
        ("distance_weight_threshold", C.c_float),
        ("distance_weight_threshold_falloff", C.c_float),
        ("max_sight_angle", C.c_float),
        ("max_sight_angle_to_lead_boid", C.c_float),
        ("time_to_start_steering_secs", C.c_uint64),
        ("time_to_start_steering_nanos", C.c_uint32),
        ("steering_overrides", C.c_int32),
    ]


_P = C.c_void_p
_PROTOS = {
    "fp_config_default": (C.c_int, [C.POINTER(FpConfig)]),
    "fp_flock_create": (C.c_int, [C.POINTER(_P), C.POINTER(FpConfig), C.c_uint64, _P, C.c_int]),
    "fp_flock_destroy": (C.c_int, [_P]),
    "fp_flock_set_config": (C.c_int, [_P, C.POINTER(FpConfig)]),
    "fp_flock_get_config": (C.c_int, [_P, C.POINTER(FpConfig)]),
    "fp_flock_set_method": (C.c_int, [_P, C.c_int]),
    "fp_flock_get_method": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "fp_flock_set_numerics": (C.c_int, [_P, C.c_int]),
    "fp_flock_get_numerics": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "fp_flock_set_leads": (C.c_int, [_P, C.c_uint32, _P]),
    "fp_flock_set_attractors": (C.c_int, [_P, C.c_uint32, _P]),
    "fp_flock_set_obstacles": (C.c_int, [_P, C.c_uint32, _P]),
    "fp_flock_set_bbox": (C.c_int, [_P, _P]),
    "fp_flock_set_lead_table": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P]),
    "fp_flock_step": (C.c_int, [_P, C.c_uint32]),
    "fp_flock_sync": (C.c_int, [_P]),
    "fp_flock_read_state": (C.c_int, [_P, _P]),
    "fp_flock_write_state": (C.c_int, [_P, _P]),
    "fp_flock_len": (C.c_uint64, [_P]),
    "fp_flock_status": (C.c_int, [_P, C.POINTER(C.c_uint32)]),
    "fp_flock_read_instances": (C.c_int, [_P, _P]),
    "fp_flock_read_instances_raw": (C.c_int, [_P, _P]),
    "fp_flock_export_instances": (C.c_int, [_P, _P, C.c_int]),
    "fp_flock_read_accel": (C.c_int, [_P, _P, _P]),
    "fp_flock_read_neighbors": (C.c_int, [_P, _P, _P]),
    "fp_flock_pair_census": (C.c_int, [_P, _P]),
    "fp_flock_set_grid_domain": (C.c_int, [_P, _P, _P]),
    "fp_flock_grid_info": (C.c_int, [_P, _P, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]),
    "fp_flock_shard_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "fp_flock_set_rebin": (C.c_int, [_P, C.c_float, C.c_float]),
    "fp_flock_rebin_info": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_uint64)]),
    "fp_flock_device_state": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P)]),
    "fp_flock_timing_begin": (C.c_int, [_P]),
    "fp_flock_timing_end": (C.c_int, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_float),
                                      C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "fp_flock_state_euler": (C.c_int, [_P, C.c_float]),
    "fp_flock_state_rk4": (C.c_int, [_P, C.c_float]),
    "fp_state_euler_combine": (C.c_int, [C.c_int, C.c_size_t, _P, _P, C.c_float, _P]),
    "fp_state_rk4_combine": (C.c_int, [C.c_int, C.c_size_t, _P, _P, _P, _P, _P, C.c_float, _P]),
    "fp_state_num_state_elements": (C.c_int, [C.c_int]),
    "fp_state_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_uint64, _P]),
    "fp_state_destroy": (C.c_int, [_P]),
    "fp_state_len": (C.c_uint64, [_P]),
    "fp_state_write": (C.c_int, [_P, _P]),
    "fp_state_read": (C.c_int, [_P, _P]),
    "fp_state_derivative": (C.c_int, [_P, _P]),
    "fp_state_euler_step": (C.c_int, [_P, C.c_float, C.c_uint32]),
    "fp_state_rk4_step": (C.c_int, [_P, C.c_float, C.c_uint32]),
    "fp_state_sync": (C.c_int, [_P]),
    "fp_state_device_vector": (C.c_int, [_P, C.POINTER(_P)]),
    "fp_state_time_steps": (C.c_int, [_P, C.c_float, C.c_int, C.c_uint32, C.POINTER(C.c_float)]),
    "fp_sph_neighbors": (C.c_int, [C.c_int, C.c_uint64, _P, C.c_uint32, C.c_float, C.c_float, _P, _P, _P,
                                   C.POINTER(C.c_float)]),
    "fp_nccl_unique_id": (C.c_int, [_P]),
    "fp_flock_create_sharded": (C.c_int, [C.POINTER(_P), C.POINTER(FpConfig), C.c_uint64, C.c_uint64,
                                          C.c_uint64, _P, C.c_int, C.c_int, C.c_int, _P]),
    "fp_flock_local_len": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "fp_flock_read_local": (C.c_int, [_P, _P, _P]),
    "fp_flock_write_local": (C.c_int, [_P, C.c_uint64, _P, _P]),
    "fp_debug_fastmath_check": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, _P]),
    "fp_last_error": (C.c_char_p, []),
    "fp_version": (C.c_char_p, []),
    "fp_launch_count": (C.c_uint64, []),
}

ABI_SYMBOLS = tuple(_PROTOS)

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FeriphysError(-2, f"{LIB_PATH} is missing: build it with "
                                    "`python -c 'import __graft_entry__ as g; g.build()'` "
                                    "(make -C feriphys_b200/csrc). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        raise FeriphysError(rc, (load().fp_last_error() or b"").decode())


def ptr(a):
    if a is None:
        return None
    return C.c_void_p(a.ctypes.data)  # (half the cost of .ctypes.data_as; the caller keeps `a` alive)


def f32c(a, shape=None) -> np.ndarray:
    out = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        out = out.reshape(shape)
    return out
