"""Frozen fixtures (tests/golden/flocking_golden.npz, made by tests/golden/make_golden.py).

The reference holds no golden vectors for flocking and cannot be run here, so the fixtures
are frozen oracle outputs: the CPU test pins the oracle to them bit for bit, the GPU tests
hold the CUDA kernels to the same immutable numbers through the C ABI."""
import os

import numpy as np
import pytest

from feriphys_b200 import synth
from oracle_lib import Scene

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "flocking_golden.npz"))
f32 = np.float32
TABLES = dict(
    attractors=np.array([[20, 20, 20, 5], [80, 10, 20, -4]], f32),
    obstacles=np.array([[40, 40, 40, 6], [10, 70, 30, 3]], f32),
    bbox=np.array([-50, 200, -50, 200, -50, 200], f32),
    leads=np.array([[5, 5, 5, 1, 0, 0, 10], [60, 9, 20, 0, 1, 0.5, 10]], f32),
)
NARROW = dict(max_sight_angle=0.8, distance_weight_threshold=5.0, distance_weight_threshold_falloff=2.5,
              centering_factor=0.3, dt=0.002)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_fixture_inputs_are_the_seeded_generator():
    assert np.array_equal(G["flock_state0"], synth.uniform_flock(600, 45.0, seed=1234))
    assert np.array_equal(G["demo1_state0"], synth.spawn_flock(synth.DEMO_SIM1["spawn"], 110))


def test_oracle_reproduces_golden(orc):
    cfg = orc.default_config()
    sc = Scene(**TABLES)
    st = G["flock_state0"]
    total, comp, flags = orc.accel_rows(cfg, sc, st)
    assert np.array_equal(bits(total), bits(G["flock_accel"]))
    assert np.array_equal(bits(comp), bits(G["flock_comp"])) and np.array_equal(flags, G["flock_flags"])
    cnt, hsh, _ = orc.neighbors_rows(cfg, st)
    assert np.array_equal(cnt, G["flock_ncount"]) and np.array_equal(hsh, G["flock_nhash"])
    cur = st
    for _ in range(10):
        cur, _ = orc.step(cfg, sc, cur)
    assert np.array_equal(bits(cur), bits(G["flock_state10"]))
    cfg2 = orc.default_config(**NARROW)
    t2, _, _ = orc.accel_rows(cfg2, None, G["narrow_state0"])
    assert np.array_equal(bits(t2), bits(G["narrow_accel"]))
    for which, scene in ((1, synth.DEMO_SIM1), (2, synth.DEMO_SIM2)):
        kinds = list(scene["lead_paths"])
        leads, times = orc.make_leads(kinds)
        cur = G[f"demo{which}_state0"]
        for _ in range(200):
            cur, _ = orc.step(cfg, Scene(leads=leads, obstacles=synth.DEMO_OBSTACLES), cur)
            orc.step_leads(leads, times, kinds, cfg.dt)
        assert np.array_equal(bits(cur), bits(G[f"demo{which}_state200"]))
        assert np.array_equal(bits(leads), bits(G[f"demo{which}_leads200"]))


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["allpairs", "grid"])
def test_cuda_matches_golden_flock(orc, method):
    from feriphys_b200 import _lib
    from gpu_util import make_pair, rel_err
    m = {"allpairs": _lib.METHOD_ALLPAIRS, "grid": _lib.METHOD_GRID}[method]
    sim, _ = make_pair(orc.default_config(), G["flock_state0"], m, TABLES)
    cnt, hsh = sim.read_neighbors()
    assert np.array_equal(cnt, G["flock_ncount"]) and np.array_equal(hsh, G["flock_nhash"])
    acc, comp = sim.read_accel(components=True)
    assert np.array_equal(bits(comp[:, 1:]), bits(G["flock_comp"][:, 1:]))
    sim.step_many(10)
    got = sim.read_state()
    if method == "allpairs":      # reference summation order: bit-identical
        assert np.array_equal(bits(acc), bits(G["flock_accel"]))
        assert np.array_equal(bits(got), bits(G["flock_state10"]))
    else:                         # cell order: 1e-5 relative (north star), 1e-4 on the trajectory
        assert rel_err(acc, G["flock_accel"]) <= 1e-5
        assert np.abs(got - G["flock_state10"]).max() <= 1e-4 * max(1.0, float(np.abs(G["flock_state10"]).max()))
    sim2, _ = make_pair(orc.default_config(**NARROW), G["narrow_state0"], m)
    c2, h2 = sim2.read_neighbors()
    assert np.array_equal(c2, G["narrow_ncount"]) and np.array_equal(h2, G["narrow_nhash"])
    a2 = sim2.read_accel()
    if method == "allpairs":
        assert np.array_equal(bits(a2), bits(G["narrow_accel"]))
    else:
        assert rel_err(a2, G["narrow_accel"]) <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("which", [1, 2])
def test_cuda_matches_golden_demo_scene(which):
    from feriphys_b200.flocking import demo_simulation
    sim = demo_simulation(which)
    assert np.array_equal(bits(sim.read_state()), bits(G[f"demo{which}_state0"]))
    sim.step_many(120)
    for _ in range(80):
        sim.step()
    assert np.array_equal(bits(sim.read_state()), bits(G[f"demo{which}_state200"]))
    assert np.array_equal(bits(np.stack([l.row() for l in sim.lead_boids])), bits(G[f"demo{which}_leads200"]))
