"""-m gpu parity of FAST numerics (fp_flock_set_numerics) against the oracle, through the C ABI.

FAST keeps the neighbour-set predicates bit-exact -- the squared distance is the reference's own,
the sight-angle decision is taken on a fused cosine only outside a 1e-5 guard band and by the
exact sequence inside it -- and evaluates the forces with FMA / MUFU.RSQ.  Bar (BASELINE.json
north_star): per-step accelerations within 1e-5 relative, 100-step trajectories within
max|dp| <= 1e-4 * max(1, |p|).  A wrong predicate decision would show as an O(1) acceleration
error of the boid concerned (one neighbour too many or too few, or the weight-1 / ramp step of
boid.rs:152-161 taken on the wrong side), so the flocks below put pairs ON the thresholds."""
import os

import numpy as np
import pytest

from feriphys_b200 import _lib, synth
from gpu_util import TABLES, make_pair, rel_err

pytestmark = pytest.mark.gpu
f32 = np.float32
NT = os.cpu_count() or 1
ACC_RTOL = 1e-5
TRAJ_TOL = 1e-4
METHODS = {"grid": _lib.METHOD_GRID, "allpairs": _lib.METHOD_ALLPAIRS}


def _fast(c, st, method, tables=None):
    sim, sc = make_pair(c, st, METHODS[method], tables, numerics=_lib.NUMERICS_FAST)
    assert sim.numerics() == (_lib.NUMERICS_FAST, _lib.NUMERICS_FAST)
    return sim, sc


def threshold_flock(seed=5):
    """Clusters of one observer and candidates placed on every decision boundary of the pair
    function: distance 15 and 16 (+- 1 ulp), exactly abeam and a hair behind / ahead of abeam,
    dead astern, coincident, equal velocity."""
    rng = np.random.default_rng(seed)
    rows = []
    for k in range(400):
        o = (rng.random(3) * 2000).astype(f32)
        v = rng.normal(size=3).astype(f32)
        axis = k % 3
        v = np.zeros(3, f32); v[axis] = f32(1.0 + (k % 5))          # flies along an axis: exact cosines
        rows.append(np.concatenate([o, v]))
        side = np.zeros(3, f32); side[(axis + 1) % 3] = 1
        fwd = np.zeros(3, f32); fwd[axis] = 1
        for d in (15.0, np.nextafter(f32(15), f32(16)), np.nextafter(f32(15), f32(0)), 16.0,
                  np.nextafter(f32(16), f32(0)), np.nextafter(f32(16), f32(17)), 15.5, 3.0):
            for direction in (fwd, side, (fwd + side) / np.sqrt(2).astype(f32)):
                p = (o + f32(d) * direction.astype(f32)).astype(f32)
                rows.append(np.concatenate([p, rng.normal(size=3).astype(f32)]))
        for eps in (0.0, 1e-7, -1e-7, 3e-6, -3e-6, 2e-5, -2e-5):            # around abeam (cos = 0 = cstar)
            p = (o + f32(7.0) * side + f32(eps * 7.0) * fwd).astype(f32)
            rows.append(np.concatenate([p, rng.normal(size=3).astype(f32)]))
        rows.append(np.concatenate([(o - f32(5.0) * fwd).astype(f32), v]))        # dead astern, same velocity
        rows.append(np.concatenate([o, rng.normal(size=3).astype(f32)]))          # coincident position
        rows.append(np.concatenate([(o + f32(1e-7) * side).astype(f32), v]))      # within abs_diff_eq of it
    return np.asarray(rows, f32)


@pytest.mark.parametrize("method", ["grid", "allpairs"])
def test_fast_accelerations_on_the_decision_boundaries(orc, method):
    st = threshold_flock()
    c = orc.default_config()
    sim, sc = _fast(c, st, method)
    ref, comp, _ = orc.accel_rows(c, sc, st, threads=NT, grid=True)
    got, gcomp = sim.read_accel(components=True)
    assert rel_err(gcomp[:, 0], comp[:, 0]) <= ACC_RTOL
    assert rel_err(got, ref) <= ACC_RTOL
    # the neighbour sets (exact kernels) agree as ever
    rc, rh, _ = orc.neighbors_rows(c, st, threads=NT, grid=True)
    gc, gh = sim.read_neighbors()
    assert np.array_equal(gc, rc) and np.array_equal(gh, rh)
    # one step: the state moves by dt * (v, a) of these accelerations
    one, _ = orc.step(c, sc, st, threads=NT, grid=True)
    sim.step()
    out = sim.read_state()
    assert np.abs(out - one).max() <= 1e-6 * max(1.0, float(np.abs(one).max()))


@pytest.mark.parametrize("method,n,extent", [("grid", 20000, 200.0), ("grid", 150000, 420.0),
                                              ("allpairs", 257, 40.0), ("allpairs", 3000, 100.0),
                                              ("allpairs", 20000, 200.0)])
def test_fast_matches_oracle_with_tables_and_trajectories(orc, method, n, extent):
    st = synth.uniform_flock(n, extent, seed=90 + n % 7)
    c = orc.default_config()
    sim, sc = _fast(c, st, method, TABLES)
    ref, comp, flags = orc.accel_rows(c, sc, st, threads=NT, grid=True)
    got, gcomp = sim.read_accel(components=True)
    for k, name in enumerate(("boids", "leads", "attractors", "bbox", "steering")):
        assert rel_err(gcomp[:, k], comp[:, k]) <= ACC_RTOL, name
    assert rel_err(got, ref) <= ACC_RTOL
    steps = 100 if n <= 20000 else 30
    cur = st
    for _ in range(steps):
        cur, _ = orc.step(c, sc, cur, threads=NT, grid=True)
    sim.step_many(steps)
    out = sim.read_state()
    scale = np.maximum(1.0, np.linalg.norm(cur[:, :3], axis=1))
    assert (np.linalg.norm(out[:, :3] - cur[:, :3], axis=1) / scale).max() <= TRAJ_TOL
    if method == "grid":
        skin, nsteps, rebins, replayed = sim.rebin_info()
        assert nsteps == steps and rebins >= 2


@pytest.mark.parametrize("method", ["grid", "allpairs"])
def test_fast_other_configs(orc, method):
    # narrow FOV + short range; FOV = pi (never culls); steering overrides
    st = synth.uniform_flock(12000, 70.0, seed=96)
    for over in (dict(max_sight_angle=0.9, distance_weight_threshold=3.0, distance_weight_threshold_falloff=2.0),
                 dict(max_sight_angle=float(f32(np.pi))), dict(max_sight_angle=0.0),
                 dict(steering_overrides=1)):
        c = orc.default_config(**over)
        sim, sc = _fast(c, st, method, TABLES)
        ref, _, _ = orc.accel_rows(c, sc, st, threads=NT, grid=True)
        got = sim.read_accel()
        assert rel_err(got, ref) <= ACC_RTOL, over
        one, _ = orc.step(c, sc, st, threads=NT, grid=True)
        sim.step()
        out = sim.read_state()
        assert np.abs(out - one).max() <= 1e-6 * max(1.0, float(np.abs(one).max())), over


def test_fast_falls_back_to_exact_on_unfilterable_thresholds(orc):
    # falloff 0: threshold and reach coincide, the ramp is empty -- FAST declines, EXACT runs
    st = synth.uniform_flock(4000, 60.0, seed=97)
    c = orc.default_config(distance_weight_threshold_falloff=0.0)
    sim, sc = make_pair(c, st, _lib.METHOD_GRID, numerics=_lib.NUMERICS_FAST)
    assert sim.numerics() == (_lib.NUMERICS_FAST, _lib.NUMERICS_EXACT)
    ref, _, _ = orc.accel_rows(c, sc, st, threads=NT, grid=True)
    assert rel_err(sim.read_accel(), ref) <= ACC_RTOL


def test_switching_numerics_mid_run(orc):
    st = synth.uniform_flock(30000, 240.0, seed=98)
    c = orc.default_config()
    sim, sc = make_pair(c, st, _lib.METHOD_GRID)
    cur = st
    for numerics, k in ((_lib.NUMERICS_EXACT, 30), (_lib.NUMERICS_FAST, 40), (_lib.NUMERICS_EXACT, 30)):
        sim.set_numerics(numerics)
        sim.step_many(k)
        for _ in range(k):
            cur, _ = orc.step(c, sc, cur, threads=NT, grid=True)
    out = sim.read_state()
    scale = np.maximum(1.0, np.linalg.norm(cur[:, :3], axis=1))
    assert (np.linalg.norm(out[:, :3] - cur[:, :3], axis=1) / scale).max() <= TRAJ_TOL
