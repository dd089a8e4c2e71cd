"""-m gpu: the device-resident State<T> (fp_state_*, fp_state.cu) and the SPH neighbour pass
(fp_sph_neighbors, fp_sph.cu) against the oracle and the reference's own State golden tests
(src/simulation/state.rs:166-185 and :218-280), through the C ABI."""
import ctypes as C
import os

import numpy as np
import pytest

from feriphys_b200 import _lib
from feriphys_b200.state import State

pytestmark = pytest.mark.gpu
f32 = np.float32


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_reference_euler_golden_state_rs_166():
    # Point at the origin, v = (0, 0, 1), constant acceleration (1, -1, 0), h = 0.5
    s = State(_lib.STATEFUL_TEST_POINT, np.array([[0, 0, 0, 0, 0, 1]], f32))
    assert s.as_vector().tolist() == [0.0, 0.0, 0.0, 0.0, 0.0, 1.0]
    s.euler_step(0.5)
    assert s.as_vector().tolist() == [0.0, 0.0, 0.5, 0.5, -0.5, 1.0]           # assert_eq! in the reference


def test_reference_rk4_golden_state_rs_218():
    s = State(_lib.STATEFUL_TEST_EXAMPLEFN, np.array([[0.5, 0.0, 0.5]], f32))
    for want_y, want_t in ((1.425130208333333, 0.5), (2.640859085770477, 1.0), (4.009155464830968, 1.5),
                           (5.305471950534675, 2.0)):
        s.rk4_step(0.5)
        y, t, h = s.as_vector().tolist()
        assert abs(y - want_y) < 0.005 and t == want_t and h == 0.5


def test_springy_point_known_answer_and_oracle(orc):
    # mass 2, force (2, -4, 6), velocity (1, 0, 0): derivative = [0, v, F / m, 0 0 0]  (springy_mesh.rs:223-240)
    one = np.array([[2, 5, 6, 7, 1, 0, 0, 2, -4, 6]], f32)
    s = State(_lib.STATEFUL_SPRINGY_POINT, one)
    assert s.derivative().tolist() == [0, 1, 0, 0, 1, -2, 3, 0, 0, 0]
    s.euler_step(0.5)
    assert s.as_vector().tolist() == [2, 5.5, 6, 7, 1.5, -1, 1.5, 2, -4, 6]
    rng = np.random.default_rng(3)
    st = rng.normal(size=(70001, 10)).astype(f32)
    st[:, 0] = np.abs(st[:, 0]) + f32(0.1)
    for integ, name in ((State.euler_step, "euler"), (State.rk4_step, "rk4")):
        dev = State(_lib.STATEFUL_SPRINGY_POINT, st)
        cur = st.reshape(-1)
        for h in (0.01, 0.5):
            integ(dev, h)
            cur = getattr(orc, "state_" + name)(cur, h, "springy_point")
            assert np.array_equal(bits(dev.as_vector()), bits(cur)), (name, h)
        # many steps in one launch == the same steps one by one
        a, b = State(_lib.STATEFUL_SPRINGY_POINT, st), State(_lib.STATEFUL_SPRINGY_POINT, st)
        integ(a, 0.01, 25)
        for _ in range(25):
            integ(b, 0.01)
        assert np.array_equal(bits(a.as_vector()), bits(b.as_vector()))


def test_rigid_body_and_boid_states_match_oracle(orc):
    rng = np.random.default_rng(4)
    n = 20003
    st = rng.normal(size=(n, 29)).astype(f32)
    q = st[:, 3:7]
    st[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)        # unit rotation
    st[:, 13] = np.abs(st[:, 13]) + f32(0.5)                          # mass
    dev = State(_lib.STATEFUL_RIGIDBODY, st)
    want = np.empty(n * 29, f32)
    orc.lib.orc_deriv_rigidbody(st.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p), C.c_size_t(n * 29), None)
    assert np.array_equal(bits(dev.derivative()), bits(want))
    cur = st.reshape(-1)
    for _ in range(3):
        dev.rk4_step(0.01)
        cur = orc.state_rk4(cur, 0.01, "rigidbody")
    assert np.array_equal(bits(dev.as_vector()), bits(cur))
    # identity rotation, identity inertia: angular velocity = angular momentum, q' = 0.5 (0, w) q
    one = np.zeros((1, 29), f32)
    one[0, 6] = 1; one[0, 13] = 2; one[0, [14, 18, 22]] = 1
    one[0, 7:10] = [2, 4, 6]; one[0, 10:13] = [0.2, 0.4, 0.6]; one[0, 23:29] = [1, 2, 3, 4, 5, 6]
    d = State(_lib.STATEFUL_RIGIDBODY, one).derivative()
    assert d[:3].tolist() == [1, 2, 3] and np.allclose(d[3:7], [0.1, 0.2, 0.3, 0.0])
    assert d[7:13].tolist() == [1, 2, 3, 4, 5, 6] and not d[13:].any()
    bo = rng.normal(size=(5000, 9)).astype(f32)
    dev = State(_lib.STATEFUL_BOID, bo)
    dev.euler_step(0.001, 10)
    cur = bo.reshape(-1)
    for _ in range(10):
        cur = orc.state_euler(cur, 0.001, "boid")
    assert np.array_equal(bits(dev.as_vector()), bits(cur))


def test_state_streams_near_the_hbm_roofline():
    # 2^24 springy points: 80 B of traffic per element and step whatever the integrator
    n = 1 << 24
    dev = State(_lib.STATEFUL_SPRINGY_POINT, np.ones((n, 10), f32))
    for rk4 in (False, True):
        dev.time_steps(0.001, rk4, 3)
        ms = dev.time_steps(0.001, rk4, 10) / 10
        gbs = n * 80 / ms / 1e6
        print(f"State<Point> {'rk4' if rk4 else 'euler'}: {ms:.3f} ms/step, {gbs:.0f} GB/s")
        assert gbs > 2000


def test_sph_neighbour_pass_matches_oracle(orc):
    from feriphys_b200.state import sph_neighbors
    # the reference's own initial lattice (sph/mod.rs:65-81): 8^3 particles 0.1 apart -- full of ties
    g = np.arange(-4, 4, dtype=np.float32) * f32(0.1)
    lattice = np.array([[x, y, z] for x in g for z in g for y in g], f32)
    rng = np.random.default_rng(5)
    cloud = (rng.random((30000, 3)) * 3.0).astype(f32)
    for pos, s in ((lattice, 0.1), (lattice, 0.25), (cloud, 0.1), (cloud, 0.17)):
        gi, gc, gd = sph_neighbors(pos, k=8, kernal_max_distance=s, particle_mass=0.001)
        ri, rc, rd = orc.sph_neighbors(pos, 8, s, 0.001)
        assert np.array_equal(gc, rc)
        assert np.array_equal(gi, ri)
        assert np.array_equal(bits(gd), bits(rd))
    assert gc.max() == 8 and gc.min() >= 1          # every particle finds at least itself
    gi, gc, gd = sph_neighbors(cloud, k=20, kernal_max_distance=0.17, particle_mass=0.001)
    ri, rc, rd = orc.sph_neighbors(cloud, 20, 0.17, 0.001)
    assert np.array_equal(gi, ri) and np.array_equal(bits(gd), bits(rd))
