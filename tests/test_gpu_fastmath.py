"""-m gpu: the branch-free sqrt / division sequences used by the grid walk's force phase
(fp_device.cuh: sqrt_rn_fast, div_rn_fast) return the IEEE round-to-nearest result --
compared bit for bit with __fsqrt_rn / __fdiv_rn on 2e9 operand sets from the admitted ranges."""
import ctypes as C

import numpy as np
import pytest

from feriphys_b200 import _lib

pytestmark = pytest.mark.gpu


def test_fast_exact_sqrt_and_div_are_correctly_rounded():
    lib = _lib.load()
    out = np.zeros(2, np.uint64)
    for seed in (1, 0xFE21F):
        _lib.check(lib.fp_debug_fastmath_check(0, 1 << 30, seed, _lib.ptr(out)))
        assert list(out) == [0, 0], out
