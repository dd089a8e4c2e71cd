"""A second, independent restatement of the reference's pair function and per-boid sums, checked
against the C oracle bit for bit on bulk random input.

The reference holds no flocking tests and cannot be built here, so the C oracle is
"parity unpinned".  What can be done about transcription errors is a second opinion: this file
restates FlockingBoid::get_acceleration and the sums of flocking.rs in vectorised numpy float32,
written from the Rust source alone (expression by expression, citing it), shares no code with
oracle/flock_oracle.c, and must agree with it in every bit -- on uniform flocks, on pairs placed
on the decision boundaries (distance 15 / 16, abeam, dead astern), on coincident positions and
equal velocities, with leads, attractors and a bounding box.  The semantics of the un-vendored
dependencies are the ones DESIGN.md section 2 declares (cgmath dot = (xx' + yy') + zz',
normalize = v * (1 / |v|), approx abs_diff_eq with f32::EPSILON, powf(2.0) = x * x,
f32::acos = this platform's libm acosf)."""
import ctypes
import ctypes.util

import numpy as np
import pytest

from oracle_lib import Scene

f32 = np.float32
EPS = np.finfo(f32).eps
_libm = ctypes.CDLL(ctypes.util.find_library("m"))
_libm.acosf.restype = ctypes.c_float
_libm.acosf.argtypes = [ctypes.c_float]


def acosf(c):
    """f32::acos -> libm acosf, element by element (numpy's arccos is another implementation)."""
    return np.array([_libm.acosf(float(x)) for x in np.asarray(c, f32).ravel()], f32).reshape(np.shape(c))


def dot(a, b):          # cgmath InnerSpace::dot for Vector3: mul_element_wise(..).sum() = (x + y) + z
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def magnitude(a):       # InnerSpace::magnitude = sqrt(dot(self, self))
    return np.sqrt(dot(a, a))


def normalize(a):       # InnerSpace::normalize = self * (1 / magnitude)
    return a * (f32(1.0) / magnitude(a))[..., None]


def abs_diff_eq(a, b):  # approx: every component (if a > b {a - b} else {b - a}) <= f32::EPSILON
    d = np.where(a > b, a - b, b - a)
    return (d <= EPS).all(axis=-1)


def get_acceleration(sp, sv, op, ov, ow, f_a, f_c, f_v, thr, fall, max_angle):
    """boid.rs:139-166 for arrays of pairs: self (sp, sv), other (op, ov, weight ow)."""
    with np.errstate(all="ignore"):
        d = op - sp                                                # other.position() - self.position
        dist = magnitude(d)                                        # distance(), boid.rs:94-96
        angle = acosf(dot(normalize(sv), normalize(d)))            # sight_angle(), boid.rs:101-107
        culled = angle > f32(max_angle)                            # boid.rs:149 (NaN: not culled)
        same_p = abs_diff_eq(op, sp)
        dhat = normalize(d)
        av = ((f32(-1.0) * f32(f_a) / (dist * dist))[..., None] * dhat) * ow[..., None]   # boid.rs:114-116
        ce = ((f32(f_c) * dist)[..., None] * dhat) * ow[..., None]                        # boid.rs:124-127
        av = np.where(same_p[..., None], f32(0), av)
        ce = np.where(same_p[..., None], f32(0), ce)
        vm = (f32(f_v) * (ov - sv)) * ow[..., None]                                       # boid.rs:135
        vm = np.where(abs_diff_eq(ov, sv)[..., None], f32(0), vm)
        w = np.where(dist <= f32(thr), f32(1.0),
                     np.where(dist >= f32(thr) + f32(fall), f32(0.0), (dist - f32(thr)) / f32(fall)))
        out = w[..., None] * ((av + ce) + vm)                                             # boid.rs:162-165
        return np.where(culled[..., None], f32(0), out).astype(f32)


def accel_from_boids(cfg, st):
    """flocking.rs:133-151: for each boid the sequential f32 sum over all others, in index order,
    skipping records equal to it (derive(PartialEq): position, velocity, weight, mass)."""
    n = len(st)
    p, v = st[:, :3], st[:, 3:]
    total = np.zeros((n, 3), f32)
    ones = np.ones(n, f32)
    for j in range(n):
        contrib = get_acceleration(p, v, np.broadcast_to(p[j], p.shape), np.broadcast_to(v[j], v.shape), ones,
                                   cfg.avoidance_factor, cfg.centering_factor, cfg.velocity_matching_factor,
                                   cfg.distance_weight_threshold, cfg.distance_weight_threshold_falloff,
                                   cfg.max_sight_angle)
        equal = (p == p[j]).all(axis=1) & (v == v[j]).all(axis=1)
        total = np.where(equal[:, None], total, total + contrib)
    return total


def accel_from_leads(cfg, st, leads):        # flocking.rs:153-170
    total = np.zeros((len(st), 3), f32)
    for l in np.asarray(leads, f32):
        n = len(st)
        total = total + get_acceleration(
            st[:, :3], st[:, 3:], np.broadcast_to(l[:3], (n, 3)), np.broadcast_to(l[3:6], (n, 3)),
            np.full(n, l[6], f32), cfg.avoidance_factor, cfg.centering_factor, cfg.velocity_matching_factor,
            cfg.distance_weight_threshold, cfg.distance_weight_threshold_falloff, cfg.max_sight_angle_to_lead_boid)
    return total


def accel_from_attractors(st, attractors):   # flocking.rs:172-180, point_attractor.rs:16-19, boid mass 1.0
    total = np.zeros((len(st), 3), f32)
    for a in np.asarray(attractors, f32):
        r = st[:, :3] - a[:3]
        mag = magnitude(r)
        total = total + ((f32(-9.8) * (a[3] + f32(1.0))) / (mag * mag))[:, None] * normalize(r)
    return total


def accel_from_bbox(st, b):                  # bounding_box.rs:13-26
    b = np.asarray(b, f32)
    x, y, z = st[:, 0], st[:, 1], st[:, 2]
    one = f32(1.0)
    return np.stack([-one / ((b[1] - x) * (b[1] - x)) + one / ((b[0] - x) * (b[0] - x)),
                     -one / ((b[3] - y) * (b[3] - y)) + one / ((b[2] - y) * (b[2] - y)),
                     one / ((b[4] - z) * (b[4] - z)) + -one / ((b[5] - z) * (b[5] - z))], axis=1).astype(f32)


def bits(a):
    return np.ascontiguousarray(a, f32).view(np.uint32)


def test_pair_function_on_the_decision_boundaries(orc):
    cfg = orc.default_config()
    rng = np.random.default_rng(11)
    sp = rng.uniform(-5, 5, (4000, 3)).astype(f32)
    sv = rng.uniform(-1, 1, (4000, 3)).astype(f32)
    dirs = rng.normal(size=(4000, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    # distances on and around 15 and 16, directions abeam / astern / ahead of the velocity
    r = rng.choice([15.0, 16.0, 15.5, 14.999999, 15.000001, 15.999999, 16.000001, 1e-3, 3.0], 4000)
    vh = sv / np.linalg.norm(sv, axis=1, keepdims=True)
    abeam = np.cross(vh, dirs)
    abeam /= np.linalg.norm(abeam, axis=1, keepdims=True)
    kind = rng.integers(0, 4, 4000)
    d = np.where((kind == 0)[:, None], dirs, np.where((kind == 1)[:, None], abeam,
                 np.where((kind == 2)[:, None], -vh, vh)))
    op = (sp + (r[:, None] * d)).astype(f32)
    ov = np.where((rng.random(4000) < 0.1)[:, None], sv, rng.uniform(-1, 1, (4000, 3))).astype(f32)
    op[:50] = sp[:50]                                     # coincident boids
    op[50:100] = sp[50:100] + f32(5e-8)                   # within f32::EPSILON
    sv[100:120] = 0                                       # zero velocity: NaN sight angle, not culled
    ours = get_acceleration(sp, sv, op, ov, np.ones(4000, f32), cfg.avoidance_factor, cfg.centering_factor,
                            cfg.velocity_matching_factor, cfg.distance_weight_threshold,
                            cfg.distance_weight_threshold_falloff, cfg.max_sight_angle)
    theirs = np.stack([orc.pair_accel(np.concatenate([sp[k], sv[k]]), op[k], ov[k]) for k in range(4000)])
    same = (bits(ours) == bits(theirs)) | (np.isnan(ours) & np.isnan(theirs))
    assert same.all(), f"{(~same).any(axis=1).sum()} pairs differ, first: {np.nonzero((~same).any(axis=1))[0][:5]}"
    assert (ours != 0).any(axis=1).sum() > 500 and (ours == 0).all(axis=1).sum() > 500


@pytest.mark.parametrize("n,extent,seed", [(300, 30.0, 1), (400, 60.0, 2)])
def test_per_boid_sums_agree_bit_for_bit(orc, n, extent, seed):
    cfg = orc.default_config()
    rng = np.random.default_rng(seed)
    st = np.concatenate([rng.uniform(0, extent, (n, 3)), rng.uniform(-1, 1, (n, 3))], axis=1).astype(f32)
    st[7] = st[3]                                          # two identical records skip each other (F8)
    st[11, :3] = st[12, :3]                                # same place, different velocity
    leads = [[5, 5, 5, 1, 0, 0, 10], [20, 9, 2, 0, 1, 0, 10]]
    attractors = [[3, 3, 3, 5], [18, 1, 2, -4]]
    bbox = [-1, extent + 1, -1, extent + 1, -1, extent + 1]
    total, comp, _ = orc.accel_rows(cfg, Scene(leads=leads, attractors=attractors, bbox=bbox), st)
    assert np.array_equal(bits(accel_from_boids(cfg, st)), bits(comp[:, 0]))
    assert np.array_equal(bits(accel_from_leads(cfg, st, leads)), bits(comp[:, 1]))
    assert np.array_equal(bits(accel_from_attractors(st, attractors)), bits(comp[:, 2]))
    assert np.array_equal(bits(accel_from_bbox(st, bbox)), bits(comp[:, 3]))
    assert (comp[:, 0] != 0).any()


# ---- steering: obstacle.rs:16-82, flocking.rs:182-209, scalar float32 with exact integer Durations ----
from fractions import Fraction  # noqa: E402

DURATION_MAX = (2 ** 64 - 1, 999_999_999)


def v_dot(a, b):
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def v_mag(a):
    return np.sqrt(v_dot(a, a))


def v_norm(a):
    return a * (f32(1.0) / v_mag(a))


def duration_from_secs_f32(x):
    """(secs, nanos): the exact value of the float in nanoseconds, rounded to nearest, ties to even
    (the declared semantics; negative / non-finite / overflowing input panics in Rust)."""
    q = Fraction(float(x)) * 10 ** 9
    ns = q.numerator // q.denominator
    rem = q - ns
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and ns % 2):
        ns += 1
    return divmod(ns, 10 ** 9)


def time_to_plane_collision(o, p, v):
    op, radius = o[:3], o[3]
    normal = v_norm(p - op)                                   # will_collide_with_plane, obstacle.rs:63-73
    denom = v_dot(normal, v)
    if not (abs(denom) > EPS):
        return None
    t = v_dot(op - p, normal) / denom
    if np.signbit(t):                                         # Signed::is_positive for floats: sign bit clear
        return None
    direction = v_norm(op - p)                                # get_velocity_components, obstacle.rs:76-81
    velocity_i = v_dot(direction, v) * direction
    return duration_from_secs_f32((v_mag(op - p) - radius) / v_mag(velocity_i))   # obstacle.rs:24-27


def steering(cfg, obstacles, p, v):
    best = None                                               # min_by keeps the first minimum
    for o in obstacles:
        t = time_to_plane_collision(o, p, v) or DURATION_MAX
        if best is None or t < best[0]:
            best = (t, o)
    t = time_to_plane_collision(best[1], p, v)
    if t is None or not t < (cfg.time_to_start_steering_secs, cfg.time_to_start_steering_nanos):
        return np.zeros(3, f32)
    o = best[1]                                               # get_acceleration_to_avoid, obstacle.rs:31-46
    direction = v_norm(o[:3] - p)
    velocity_t = v - v_dot(direction, v) * direction
    ts = f32(t[0]) + f32(t[1]) / f32(1_000_000_000)           # Duration::as_secs_f32
    if ts * v_mag(velocity_t) > o[3]:
        return np.zeros(3, f32)
    return (f32(2.0) * (o[3] - ts * v_mag(velocity_t)) / (ts * ts)) * v_norm(velocity_t)


def test_steering_agrees_bit_for_bit(orc):
    cfg = orc.default_config()
    rng = np.random.default_rng(5)
    obstacles = np.array([[40, 40, 40, 6], [10, 70, 30, 3], [70, 20, 60, 9], [40, 40, 40, 6]], f32)  # (a tie)
    sc = Scene(obstacles=obstacles)
    n_steer = 0
    for _ in range(1500):
        p = rng.uniform(0, 80, 3).astype(f32)
        if min(np.linalg.norm(p - o[:3]) - o[3] for o in obstacles) < 0.5:
            continue                                          # inside a sphere the reference panics (F10)
        k = rng.integers(0, 4)
        aim = obstacles[k, :3] + rng.normal(size=3).astype(f32) * obstacles[k, 3]
        v = ((aim - p) / np.linalg.norm(aim - p) * rng.uniform(0.5, 30)).astype(f32) if rng.random() < 0.8 \
            else rng.uniform(-1, 1, 3).astype(f32)
        with np.errstate(all="ignore"):
            ours = steering(cfg, obstacles, p, v)
        theirs, flags = orc.steering_accel(cfg, sc, np.concatenate([p, v]))
        assert flags == 0
        assert np.array_equal(bits(ours), bits(theirs)), (p, v, ours, theirs)
        n_steer += bool((ours != 0).any())
    assert n_steer > 100


def test_whole_step_agrees_bit_for_bit(orc):
    """Simulation::step (flocking.rs:97-122): a = (((boids + leads) + attractors) + bbox) + steering,
    or steering alone when it overrides; p' = p + dt * v, v' = v + dt * a, from the old state only."""
    rng = np.random.default_rng(9)
    n, extent = 250, 40.0
    st = np.concatenate([rng.uniform(0, extent, (n, 3)), rng.uniform(-3, 3, (n, 3))], axis=1).astype(f32)
    leads = [[5, 5, 5, 1, 0, 0, 10], [20, 9, 2, 0, 1, 0, 10]]
    attractors = [[3, 3, 3, 5], [18, 1, 2, -4]]
    obstacles = np.array([[60, 20, 20, 6], [-25, 20, 20, 9]], f32)      # outside the flock: nobody is inside a sphere
    for i in range(0, n, 5):                                            # every fifth boid heads for one, fast
        aim = obstacles[i % 2, :3] + rng.normal(size=3).astype(f32) * f32(3)
        st[i, 3:] = (aim - st[i, :3]) / np.linalg.norm(aim - st[i, :3]) * f32(25)
    bbox = [-1, extent + 1, -1, extent + 1, -1, extent + 1]
    sc = Scene(leads=leads, attractors=attractors, obstacles=obstacles, bbox=bbox)
    for overrides in (0, 1):
        cfg = orc.default_config(steering_overrides=overrides)
        with np.errstate(all="ignore"):
            steer = np.stack([steering(cfg, obstacles, st[i, :3], st[i, 3:]) for i in range(n)]).astype(f32)
            if overrides:
                a = steer
            else:
                a = (((accel_from_boids(cfg, st) + accel_from_leads(cfg, st, leads)) + accel_from_attractors(st, attractors))
                     + accel_from_bbox(st, bbox)) + steer
        dt = f32(cfg.dt)
        exp = np.concatenate([st[:, :3] + dt * st[:, 3:], st[:, 3:] + dt * a], axis=1).astype(f32)
        got, flags = orc.step(cfg, sc, st)
        assert not flags.any()
        assert np.array_equal(bits(exp), bits(got))
        assert (steer != 0).any()
