"""Known answers for the oracle.

The reference has no flocking tests, so these are hand-derived
from the Rust source: boid.rs:94-166, point_attractor.rs:16-19,
bounding_box.rs:13-26, obstacle.rs, flocking.rs:182-209.  Each expected value
is computed here in numpy float32 following the source expression, or is an
exactly representable number.
"""
import math

import numpy as np
import pytest

from oracle_lib import FLAG_STEER_NEGATIVE, FLAG_STEER_NAN_OVF, Scene

f32 = np.float32
SELF = [0, 0, 0, 1, 0, 0]


def test_default_config_matches_flocking_rs_36_51(orc):
    c = orc.default_config()
    assert c.dt == f32(1_000_000) / f32(1e9)  # Duration::from_millis(1).as_secs_f32()
    assert c.dt == f32(0.001)
    assert (c.avoidance_factor, c.centering_factor, c.velocity_matching_factor) == (
        f32(1.0), f32(0.1), f32(0.5))
    assert (c.distance_weight_threshold, c.distance_weight_threshold_falloff) == (15.0, 1.0)
    assert c.max_sight_angle == f32(np.pi) / f32(2)
    assert c.max_sight_angle_to_lead_boid == f32(np.pi)
    assert (c.time_to_start_steering_secs, c.time_to_start_steering_nanos) == (4, 0)
    assert c.steering_overrides == 0


def test_b1_ahead_in_range(orc):
    # av = -1/4 * (1,0,0); ce = 0.1*2 * (1,0,0); vm = 0 (equal velocities)
    a = orc.pair_accel(SELF, [2, 0, 0], [1, 0, 0])
    exp = f32(-1.0) / f32(4.0) + f32(0.1) * f32(2.0)
    assert a[0] == exp and a[1] == 0 and a[2] == 0
    assert abs(float(a[0]) + 0.05) < 1e-7


def test_b2_directly_behind_is_culled(orc):
    assert orc.sight_angle(SELF, [-2, 0, 0]) == f32(np.pi)
    assert np.all(orc.pair_accel(SELF, [-2, 0, 0], [5, 5, 5]) == 0)


def test_b3_exactly_abeam_is_kept(orc):
    # acos(0) = 1.57079637 == max_sight_angle; the test is '>' (boid.rs:149)
    assert orc.sight_angle(SELF, [0, 2, 0]) == f32(np.pi) / f32(2)
    a = orc.pair_accel(SELF, [0, 2, 0], [1, 0, 0])
    assert a[0] == 0 and a[2] == 0
    assert a[1] == f32(-0.25) + f32(0.1) * f32(2.0)


def test_b4_ramp_is_inverted_F7(orc):
    # weight (d - thr)/fall: 0.5 at 15.5, 1.0 at <= 15, 0 at >= 16 (boid.rs:152-161)
    def w(d):
        a = orc.pair_accel(SELF, [d, 0, 0], [1, 0, 0])
        full = f32(-1.0) / (f32(d) * f32(d)) + f32(0.1) * f32(d)
        return a[0], full
    a, full = w(15.5)
    assert a == f32(0.5) * full
    a, full = w(15.0)
    assert a == full
    a, full = w(16.0)
    assert a == 0
    a, full = w(15.0 + 2 ** -10)   # just past thr: weight is ~0, not ~1
    assert abs(a) < abs(full) * 1e-3


def test_b5_velocity_matching(orc):
    a = orc.pair_accel(SELF, [2, 0, 0], [1, 1, 0])
    assert a[1] == f32(0.5) and a[2] == 0
    assert a[0] == f32(-0.25) + f32(0.1) * f32(2.0)


def test_b6_lead_weight(orc):
    a = orc.pair_accel(SELF, [2, 0, 0], [1, 0, 0], weight=10.0, theta=float(f32(np.pi)))
    assert a[0] == f32(-0.25) * f32(10) + (f32(0.1) * f32(2.0)) * f32(10)
    # behind, but leads use max_sight_angle_to_lead_boid = pi: acos(-1) == pi, not '>'
    a = orc.pair_accel(SELF, [-2, 0, 0], [1, 0, 0], weight=10.0, theta=float(f32(np.pi)))
    assert a[0] != 0


def test_b7_duplicates_skip_each_other(orc):
    st = np.array([[1, 2, 3, 1, 0, 0], [1, 2, 3, 1, 0, 0], [3, 2, 3, 1, 0, 0]], f32)
    cfg = orc.default_config()
    total, comp, _ = orc.accel_rows(cfg, None, st)
    # rows 0 and 1 only see boid 2; identical results
    assert np.array_equal(total[0], total[1])
    one = orc.pair_accel(st[0], st[2, :3], st[2, 3:])
    assert np.array_equal(total[0], one)
    cnt, _, lst = orc.neighbors_rows(cfg, st, list_cap=4)
    assert list(cnt) == [1, 1, 0]  # boid 2 looks along +x; 0 and 1 are behind it
    assert lst[0, 0] == 2 and lst[1, 0] == 2


def test_b8_zero_velocity_is_not_culled(orc):
    # normalize(0) = 0 * (1/0) = NaN; acos(NaN) > theta is false => forces evaluated
    a = orc.pair_accel([0, 0, 0, 0, 0, 0], [-2, 0, 0], [0, 0, 0])
    assert math.isnan(orc.sight_angle([0, 0, 0, 0, 0, 0], [-2, 0, 0]))
    assert a[0] == -(f32(-0.25) + f32(0.1) * f32(2.0))


def test_b9_attractor(orc):
    a = orc.attractor_accel([0, 0, 0, 9], [2, 0, 0])
    assert a[0] == (f32(-9.8) * f32(10)) / f32(4) and a[1] == 0 and a[2] == 0
    assert abs(float(a[0]) + 24.5) < 1e-5
    # negative mass repels
    a = orc.attractor_accel([0, 0, 0, -11], [2, 0, 0])
    assert a[0] > 0


def test_b10_bbox(orc):
    a = orc.bbox_accel([0, 10, 0, 10, 0, 10], [2, 5, 5])
    assert a[0] == f32(-1 / 64) + f32(0.25) and a[1] == 0 and a[2] == 0
    # z adds back (start) first then front (end): bounding_box.rs:25
    a = orc.bbox_accel([0, 10, 0, 10, 0, 10], [5, 5, 1])
    assert a[2] == f32(1.0) + f32(-1.0) / f32(81.0)


def test_b11_steering_gate(orc):
    cfg = orc.default_config()
    sc = Scene(obstacles=[[10, 0, 0, 4]])
    a, fl = orc.steering_accel(cfg, sc, [0, 0, 0, 1, 0.1, 0])
    assert fl == 0 and np.all(a == 0)          # 6 s away >= 4 s
    a, fl = orc.steering_accel(cfg, sc, [3, 0, 0, 1, 0.1, 0])
    assert fl == 0
    # t = 3 s exactly; |v_t| = 0.1f; 2*(4 - 3*0.1)/9 * (0,1,0)
    exp = (f32(2.0) * (f32(4.0) - f32(3.0) * f32(0.1))) / f32(9.0)
    assert a[0] == 0 and a[2] == 0
    assert abs(float(a[1]) - float(exp)) <= 2e-7 * float(exp)
    assert abs(float(a[1]) - 0.82222) < 1e-4
    # moving away: never hits the plane
    a, fl = orc.steering_accel(cfg, sc, [3, 0, 0, -1, 0.1, 0])
    assert fl == 0 and np.all(a == 0)
    # tangential slip too large: t * |v_t| > r => 0 (obstacle.rs:36-38)
    a, fl = orc.steering_accel(cfg, sc, [3, 0, 0, 1, 2.0, 0])
    assert fl == 0 and np.all(a == 0)


def test_steering_picks_first_minimum_and_flags_panics(orc):
    cfg = orc.default_config()
    # two obstacles at the same time-to-plane: min_by keeps the first
    sc = Scene(obstacles=[[10, 0, 0, 4], [10, 0, 0, 2]])
    a1, _ = orc.steering_accel(cfg, sc, [5, 0, 0, 2, 0.1, 0])      # t = (5-4)/2 = .5 ; (5-2)/2 = 1.5
    only_first, _ = orc.steering_accel(cfg, Scene(obstacles=[[10, 0, 0, 4]]), [5, 0, 0, 2, 0.1, 0])
    assert np.array_equal(a1, only_first)
    # inside the sphere and closing: Duration::from_secs_f32(negative) panics (F10)
    a, fl = orc.steering_accel(cfg, Scene(obstacles=[[10, 0, 0, 4]]), [7, 0, 0, 1, 0.1, 0])
    assert fl == FLAG_STEER_NEGATIVE and np.all(a == 0)
    # a panicking obstacle anywhere in the list poisons the boid (min_by evaluates all)
    a, fl = orc.steering_accel(cfg, Scene(obstacles=[[30, 0, 0, 4], [10, 0, 0, 4]]),
                               [7, 0, 0, 1, 0.1, 0])
    assert fl == FLAG_STEER_NEGATIVE and np.all(a == 0)


def test_duration_from_secs_f32_rounds_to_nearest_ns(orc):
    d = orc.duration_from_secs_f32
    assert d(0.0) == (0, 0, 0)
    assert d(-0.0) == (0, 0, 0)
    assert d(1.0) == (0, 1, 0)
    assert d(0.5) == (0, 0, 500_000_000)
    # f32(0.001) = 0.001000000047497451305389404296875 -> 1_000_000 ns
    assert d(float(f32(0.001))) == (0, 0, 1_000_000)
    # 2.7 as f32 = 2.7000000476837158203125
    assert d(float(f32(2.7))) == (0, 2, 700_000_048)
    # below half a nanosecond
    assert d(4e-10) == (0, 0, 0)
    assert d(6e-10) == (0, 0, 1)
    assert d(float(f32(2.0 ** 40))) == (0, 2 ** 40, 0)
    assert d(-1e-9)[0] == FLAG_STEER_NEGATIVE
    assert d(float("nan"))[0] == FLAG_STEER_NAN_OVF
    assert d(float(f32(2.0 ** 64)))[0] == FLAG_STEER_NAN_OVF
    assert d(float("inf"))[0] == FLAG_STEER_NAN_OVF
    assert orc.duration_as_secs_f32(2, 700_000_048) == float(f32(2) + f32(700_000_048) / f32(1e9))


def test_duration_exhaustive_sample_against_python_fractions(orc):
    from fractions import Fraction
    rng = np.random.default_rng(7)
    xs = np.concatenate([
        rng.uniform(0, 10, 2000), rng.uniform(0, 1e-6, 500), rng.uniform(1e3, 1e7, 500),
        10.0 ** rng.uniform(-12, 12, 1000)]).astype(f32)
    for x in xs:
        ns = Fraction(float(x)) * 10 ** 9
        fl = ns.numerator // ns.denominator
        rem = ns - fl
        if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and fl % 2 == 1):
            fl += 1
        f, s, n = orc.duration_from_secs_f32(float(x))
        assert f == 0 and s * 10 ** 9 + n == fl, x


def test_step_is_jacobi_euler(orc):
    # flocking.rs:116-117: p' = p + dt*v ; v' = v + dt*a with a from the OLD state
    cfg = orc.default_config()
    st = np.array([[0, 0, 0, 1, 0, 0], [2, 0, 0, 1, 1, 0], [1, 3, 0, 0, -1, 0]], f32)
    total, _, _ = orc.accel_rows(cfg, None, st)
    out, _ = orc.step(cfg, None, st)
    dt = f32(cfg.dt)
    assert np.array_equal(out[:, :3], st[:, :3] + dt * st[:, 3:])
    assert np.array_equal(out[:, 3:], st[:, 3:] + dt * total)


def test_accel_sum_order_and_components(orc):
    cfg = orc.default_config()
    rng = np.random.default_rng(3)
    st = rng.uniform(0, 10, (40, 6)).astype(f32)
    sc = Scene(leads=[[5, 5, 5, 1, 0, 0, 10], [0, 9, 2, 0, 1, 0, 10]],
               attractors=[[3, 3, 3, 5], [8, 1, 2, -4]],
               obstacles=[[12, 5, 5, 2], [5, 14, 5, 3]], bbox=[-1, 11, -1, 11, -1, 11])
    total, comp, flags = orc.accel_rows(cfg, sc, st)
    exp = (((comp[:, 0] + comp[:, 1]) + comp[:, 2]) + comp[:, 3]) + comp[:, 4]
    assert np.array_equal(total, exp)
    assert np.any(comp[:, 1] != 0) and np.any(comp[:, 2] != 0) and np.any(comp[:, 3] != 0)
    # the boid-boid component is the sequential sum of pair accelerations in index order
    i = 7
    acc = np.zeros(3, f32)
    for j in range(len(st)):
        if j == i:
            continue
        acc = acc + orc.pair_accel(st[i], st[j, :3], st[j, 3:])
    assert np.array_equal(acc, comp[i, 0])
    # steering_overrides: only the steering term (flocking.rs:102-103)
    cfg2 = orc.default_config(steering_overrides=1)
    t2, c2, _ = orc.accel_rows(cfg2, sc, st)
    assert np.array_equal(t2, c2[:, 4])
    # omp rows are identical to the single-threaded loop
    t3, _, _ = orc.accel_rows(cfg, sc, st, threads=4)
    assert np.array_equal(total, t3)


def test_acos_threshold_form_of_fov_predicate(orc):
    # culled <=> -1 <= c <= c*(theta)
    assert orc.acos_threshold(float(f32(np.pi) / f32(2))) == pytest.approx(-1.03316033e-07, rel=1e-6)
    assert np.float32(orc.acos_threshold(float(f32(np.pi) / f32(2)))).view(np.uint32) == 0xB3DDDE97
    assert orc.acos_threshold(float(f32(np.pi))) == -2.0          # never culls
    assert orc.acos_threshold(-0.5) == 1.0                        # always culls
    assert orc.acos_threshold(float("nan")) == -2.0
    for theta in (0.0, 0.3, 1.0, float(f32(np.pi) / f32(4)), 2.5, 3.1):
        cs = f32(orc.acos_threshold(theta))
        nxt = np.nextafter(cs, f32(2))
        assert math.acos(min(1.0, float(cs))) > theta - 1e-6
        self6 = [0, 0, 0, 1, 0, 0]
        # direct check of the two sides through acosf itself
        import ctypes
        libm = ctypes.CDLL("libm.so.6")
        libm.acosf.restype = ctypes.c_float
        libm.acosf.argtypes = [ctypes.c_float]
        assert libm.acosf(float(cs)) > f32(theta)
        assert not (libm.acosf(float(nxt)) > f32(theta))


def test_acosf_monotone_sampled(orc):
    # full sweep is tests/test_oracle_kat.py::test_acosf_monotone_full (slow marker)
    assert orc.acos_monotone_violations(0x3F000000, 0x3F000000 + (1 << 22)) == 0   # [0.5, ...)
    assert orc.acos_monotone_violations(0xBF800000 - (1 << 22), 0xBF800000) == 0   # up to -1
    assert orc.acos_monotone_violations(0x00000000, 1 << 22) == 0
    assert orc.acos_monotone_violations(0x33000000, 0x33000000 + (1 << 24)) == 0   # around 3e-8
    assert orc.acos_monotone_violations(0xB3000000, 0xB3000000 + (1 << 24)) == 0


@pytest.mark.slow
def test_acosf_monotone_full(orc):
    assert orc.acos_monotone_violations(0x00000000, 0x3F800000) == 0
    assert orc.acos_monotone_violations(0x80000000, 0xBF800000) == 0
