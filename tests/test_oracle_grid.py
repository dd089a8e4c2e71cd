"""The grid-accelerated oracle mode must be bit-identical to the literal
O(N^2) loops (flocking.rs:133-151): same neighbour sets, same accelerations,
same stepped state.  Also pins lead-boid stepping (boid.rs:46-53)."""
import numpy as np

from feriphys_b200 import synth
from oracle_lib import Scene

f32 = np.float32


def _cases(orc):
    yield "uniform-sparse", synth.uniform_flock(3000, 120.0, seed=11), orc.default_config()
    yield "dense-demo", synth.spawn_flock([(25, 0.5, 0)], 110, seed=5), orc.default_config()
    yield "fov-narrow", synth.uniform_flock(2000, 60.0, seed=12), orc.default_config(
        max_sight_angle=0.7, distance_weight_threshold=6.0, distance_weight_threshold_falloff=3.0)
    yield "no-falloff", synth.uniform_flock(1500, 50.0, seed=13), orc.default_config(
        distance_weight_threshold_falloff=0.0, max_sight_angle=float(f32(np.pi)))


def test_grid_mode_equals_literal_mode(orc):
    for name, st, cfg in _cases(orc):
        sc = Scene(leads=[[5, 5, 5, 1, 0, 0, 10]], attractors=[[20, 20, 20, 5]],
                   obstacles=[[40, 40, 40, 6]], bbox=[-50, 200, -50, 200, -50, 200])
        t0, c0, f0 = orc.accel_rows(cfg, sc, st, threads=8)
        t1, c1, f1 = orc.accel_rows(cfg, sc, st, threads=8, grid=True)
        assert np.array_equal(t0.view(np.uint32), t1.view(np.uint32)), name
        assert np.array_equal(c0.view(np.uint32), c1.view(np.uint32)), name
        n0, h0, _ = orc.neighbors_rows(cfg, st, threads=8)
        n1, h1, _ = orc.neighbors_rows(cfg, st, threads=8, grid=True)
        assert np.array_equal(n0, n1) and np.array_equal(h0, h1), name
        s0, _ = orc.step(cfg, sc, st, threads=8)
        s1, _ = orc.step(cfg, sc, st, threads=8, grid=True)
        assert np.array_equal(s0.view(np.uint32), s1.view(np.uint32)), name


def test_neighbor_hash_is_sum_of_mix64(orc):
    cfg = orc.default_config()
    st = synth.uniform_flock(300, 30.0, seed=2)
    cnt, hsh, lst = orc.neighbors_rows(cfg, st, list_cap=300)
    for i in (0, 17, 299):
        js = [int(j) for j in lst[i, :cnt[i]]]
        assert js == sorted(js)
        assert sum(orc.mix64(j) for j in js) % 2 ** 64 == int(hsh[i])


def test_pair_census_adds_up(orc):
    cfg = orc.default_config()
    st = synth.uniform_flock(500, 40.0, seed=4)
    c = orc.pair_census(cfg, st, threads=4)
    assert int(c.sum()) == 500 * 499
    cnt, _, _ = orc.neighbors_rows(cfg, st)
    assert int(cnt.sum()) == int(c[2])


def test_lead_step_lags_one_step_F9(orc):
    # boid.rs:46-53 / parametric.rs:17-21: returns path(t) THEN advances t
    leads, times = orc.make_leads([0])
    assert np.array_equal(leads[0, :3], [25.0, 0.5, 0.0]) and np.all(leads[0, 3:6] == 0)
    dt = float(f32(0.001))
    orc.step_leads(leads, times, [0], dt)
    # first step evaluates path(0) again => velocity exactly 0
    assert np.array_equal(leads[0, :3], [25.0, 0.5, 0.0]) and np.all(leads[0, 3:6] == 0)
    assert times[0] == f32(0.001)
    orc.step_leads(leads, times, [0], dt)
    p = orc.demo_path(0, float(f32(0.001)))
    assert np.array_equal(leads[0, :3], p)
    assert leads[0, 3] == (p[0] - f32(25.0)) / f32(0.001)
    assert times[0] == f32(0.001) + f32(0.001)
    # a linear path through the generic callback
    f, row, t = orc.lead_step_fn([0, 0, 0, 0, 0, 0, 10], 0.5, 0.25, lambda t: (t, 2 * t, 1.0))
    assert f == 0 and t == 0.75
    assert list(row[:3]) == [0.5, 1.0, 1.0] and list(row[3:6]) == [2.0, 4.0, 4.0]
    # zero Duration returns early (boid.rs:47-49); 4e-10 s rounds to 0 ns
    f, row, t = orc.lead_step_fn([1, 2, 3, 0, 0, 0, 10], 0.5, 4e-10, lambda t: (9, 9, 9))
    assert f == 0 and t == 0.5 and list(row[:3]) == [1, 2, 3]
