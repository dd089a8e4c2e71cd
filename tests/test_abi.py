"""The C-ABI library loads without a GPU and exports every symbol the header
declares; compute entry points fail loudly (no CPU fallback) when no device
exists."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from feriphys_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "feriphys_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(fp_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/feriphys_cuda.h but not exported"
    assert set(names) == set(_lib.ABI_SYMBOLS), set(names) ^ set(_lib.ABI_SYMBOLS)


def test_config_default_and_struct_layout():
    lib = _lib.load()
    c = _lib.FpConfig()
    assert lib.fp_config_default(C.byref(c)) == 0
    assert c.dt == np.float32(0.001) and c.distance_weight_threshold == 15.0
    assert c.max_sight_angle == np.float32(np.pi) / np.float32(2)
    assert c.time_to_start_steering_secs == 4 and c.steering_overrides == 0
    assert C.sizeof(_lib.FpConfig) == 48
    assert lib.fp_version().startswith(b"feriphys-cuda")


def test_python_config_matches_c_default():
    from feriphys_b200.flocking import Config
    lib = _lib.load()
    c = _lib.FpConfig()
    lib.fp_config_default(C.byref(c))
    p = Config().to_c()
    for name, _ in _lib.FpConfig._fields_:
        assert getattr(c, name) == getattr(p, name), name


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _lib.load()
    h = C.c_void_p()
    st = np.zeros((4, 6), np.float32)
    rc = lib.fp_flock_create(C.byref(h), None, 4, st.ctypes.data_as(C.c_void_p), 0)
    assert rc == -2 and b"no CPU fallback" in lib.fp_last_error()
    out = np.zeros(4, np.float32)
    rc = lib.fp_state_euler_combine(0, 4, out.ctypes.data_as(C.c_void_p),
                                    out.ctypes.data_as(C.c_void_p), 0.5,
                                    out.ctypes.data_as(C.c_void_p))
    assert rc == -2
