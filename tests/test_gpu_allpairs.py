"""-m gpu parity of the tiled all-pairs kernel (K1) and the single-CTA kernel
(K4) against the oracle, through the C ABI.

Bar: neighbour sets bit-exact; accelerations and stepped states BIT-IDENTICAL
(these kernels keep the reference's summation order and use unfused IEEE
arithmetic, so the north-star tolerance of 1e-5 relative is met with 0)."""
import os

import numpy as np
import pytest

from feriphys_b200 import _lib, synth
from feriphys_b200.flocking import Duration, demo_simulation
from gpu_util import TABLES, bits, make_pair, py_config

pytestmark = pytest.mark.gpu
f32 = np.float32
NT = os.cpu_count() or 1


def _kat_flock(orc, st, method, **cfg):
    c = orc.default_config(**cfg)
    sim, sc = make_pair(c, np.asarray(st, f32), method)
    ref, comp, _ = orc.accel_rows(c, sc, np.asarray(st, f32))
    got = sim.read_accel()
    assert np.array_equal(bits(got), bits(ref)), (got, ref)
    return got


@pytest.mark.parametrize("method", [_lib.METHOD_ALLPAIRS, _lib.METHOD_GRID])
def test_known_answers_appendix_b(orc, method):
    a = _kat_flock(orc, [[0, 0, 0, 1, 0, 0], [2, 0, 0, 1, 0, 0]], method)          # B1
    assert a[0, 0] == f32(-0.25) + f32(0.1) * f32(2)
    assert np.all(a[1] == 0)                                                        # B2: 0 is behind 1
    a = _kat_flock(orc, [[0, 0, 0, 1, 0, 0], [0, 2, 0, 1, 0, 0]], method)          # B3 abeam kept
    assert a[0, 1] == f32(-0.25) + f32(0.1) * f32(2)
    a = _kat_flock(orc, [[0, 0, 0, 1, 0, 0], [15.5, 0, 0, 1, 0, 0]], method)       # B4 ramp
    assert a[0, 0] == f32(0.5) * (f32(-1) / (f32(15.5) * f32(15.5)) + f32(0.1) * f32(15.5))
    a = _kat_flock(orc, [[0, 0, 0, 1, 0, 0], [16.0, 0, 0, 1, 0, 0]], method)
    assert np.all(a == 0)
    a = _kat_flock(orc, [[0, 0, 0, 1, 0, 0], [2, 0, 0, 1, 1, 0]], method)          # B5
    assert a[0, 1] == f32(0.5)
    a = _kat_flock(orc, [[1, 2, 3, 1, 0, 0], [1, 2, 3, 1, 0, 0], [3, 2, 3, 1, 0, 0]], method)  # B7
    assert np.array_equal(a[0], a[1])
    a = _kat_flock(orc, [[0, 0, 0, 0, 0, 0], [-2, 0, 0, 0, 0, 0]], method)         # B8 zero velocity
    assert a[0, 0] == -(f32(-0.25) + f32(0.1) * f32(2))


def test_extras_known_answers(orc):
    c = orc.default_config()
    for st, tables in [
        ([[2, 0, 0, 1, 0, 0]], dict(attractors=np.array([[0, 0, 0, 9]], f32))),                 # B9
        ([[2, 5, 5, 1, 0, 0]], dict(bbox=np.array([0, 10, 0, 10, 0, 10], f32))),                # B10
        ([[3, 0, 0, 1, 0.1, 0], [0, 0, 0, 1, 0.1, 0], [3, 0, 0, -1, 0.1, 0], [3, 0, 0, 1, 2, 0]],
         dict(obstacles=np.array([[10, 0, 0, 4]], f32))),                                        # B11
        ([[0, 0, 0, 1, 0, 0]], dict(leads=np.array([[2, 0, 0, 1, 0, 0, 10]], f32))),            # B6
    ]:
        sim, sc = make_pair(c, np.asarray(st, f32), _lib.METHOD_ALLPAIRS, tables)
        ref, comp, _ = orc.accel_rows(c, sc, np.asarray(st, f32))
        got, gcomp = sim.read_accel(components=True)
        assert np.array_equal(bits(got), bits(ref))
        assert np.array_equal(bits(gcomp), bits(comp))
        assert sim.status() == 0


def test_steering_panic_flags(orc):
    c = orc.default_config()
    st = np.array([[7, 0, 0, 1, 0.1, 0], [0, 0, 0, 1, 0.1, 0]], f32)   # boid 0 is inside the sphere
    sim, sc = make_pair(c, st, _lib.METHOD_ALLPAIRS, dict(obstacles=np.array([[10, 0, 0, 4]], f32)))
    ref, comp, flags = orc.accel_rows(c, sc, st)
    got, gcomp = sim.read_accel(components=True)
    assert flags[0] == 1 and flags[1] == 0
    assert np.array_equal(bits(gcomp[:, 4]), bits(comp[:, 4])) and np.all(gcomp[0, 4] == 0)
    assert sim.status() == _lib.STATUS_STEER_NEGATIVE
    assert sim.status() == 0   # reading clears


CASES = [
    ("uniform-2k-defaults", dict(n=2000, extent=60.0, seed=21), {}, True),
    ("dense-c2-like", dict(n=1500, extent=24.0, seed=22), dict(max_sight_angle=float(f32(np.pi))), False),
    ("narrow-fov", dict(n=1777, extent=40.0, seed=23),
     dict(max_sight_angle=0.7, distance_weight_threshold=6.0, distance_weight_threshold_falloff=3.0), True),
    ("no-falloff", dict(n=900, extent=30.0, seed=24), dict(distance_weight_threshold_falloff=0.0), False),
    ("override", dict(n=700, extent=30.0, seed=25), dict(steering_overrides=1), True),
    ("ragged-129", dict(n=129, extent=20.0, seed=26), {}, True),
    ("single", dict(n=1, extent=20.0, seed=27), {}, True),
]


@pytest.mark.parametrize("name,flock,cfg,tables", CASES, ids=[c[0] for c in CASES])
def test_allpairs_bit_exact_vs_oracle(orc, name, flock, cfg, tables):
    c = orc.default_config(**cfg)
    st = synth.uniform_flock(flock["n"], flock["extent"], seed=flock["seed"])
    sim, sc = make_pair(c, st, _lib.METHOD_ALLPAIRS, TABLES if tables else None)
    # neighbour sets: bit-exact predicate check
    rc, rh, _ = orc.neighbors_rows(c, st, threads=NT)
    gc, gh = sim.read_neighbors()
    assert np.array_equal(gc, rc) and np.array_equal(gh, rh)
    # accelerations, per component
    ref, comp, flags = orc.accel_rows(c, sc, st, threads=NT)
    got, gcomp = sim.read_accel(components=True)
    assert np.array_equal(bits(gcomp), bits(comp))
    assert np.array_equal(bits(got), bits(ref))
    # census agrees with the oracle's
    cen = sim.pair_census()
    ocen = orc.pair_census(c, st, threads=NT)
    assert list(cen[:3]) == list(ocen) and cen[3] == ocen.sum()
    # 100-step trajectory, bit-identical
    cur = st
    for _ in range(100):
        cur, _ = orc.step(c, sc, cur, threads=NT)
    sim.step_many(100)
    assert np.array_equal(bits(sim.read_state()), bits(cur))


def test_empty_flock():
    from feriphys_b200.flocking import Simulation
    sim = Simulation.from_state(np.zeros((0, 6), f32))
    sim.step()
    assert sim.read_state().shape == (0, 6)
    assert list(sim.pair_census()) == [0, 0, 0, 0]


@pytest.mark.parametrize("which", [1, 2])
@pytest.mark.parametrize("method", [_lib.METHOD_SMALL, _lib.METHOD_ALLPAIRS])
def test_demo_scene_1000_steps_bit_exact(orc, which, method):
    """Config C1: demos/flocking.rs:92-156, 110 boids, lead boids, ship obstacle,
    1000 headless steps -- the whole trajectory is bit-identical to the oracle."""
    from oracle_lib import Scene
    scene = synth.DEMO_SIM1 if which == 1 else synth.DEMO_SIM2
    kinds = list(scene["lead_paths"])
    st = synth.spawn_flock(scene["spawn"], scene["num_boids"])
    c = orc.default_config()
    leads, times = orc.make_leads(kinds)
    cur = st
    for _ in range(1000):
        sc = Scene(leads=leads, obstacles=synth.DEMO_OBSTACLES)
        cur, fl = orc.step(c, sc, cur)
        assert not fl.any()
        orc.step_leads(leads, times, kinds, c.dt)
    sim = demo_simulation(which, method=method)
    assert np.array_equal(bits(sim.read_state()), bits(st))
    if method == _lib.METHOD_SMALL:
        sim.step_many(400)
        sim.step_many(600)
    else:
        for _ in range(3):
            sim.step()              # the per-step path: upload leads, step, advance leads
        sim.step_many(997)
    assert sim.method_in_use() == method
    assert np.array_equal(bits(sim.read_state()), bits(cur))
    assert sim.status() == 0
    # lead boids ended where the oracle's did
    assert np.array_equal(bits(np.stack([l.row() for l in sim.lead_boids])), bits(leads))


def test_config_change_between_steps(orc):
    c = orc.default_config()
    st = synth.uniform_flock(800, 30.0, seed=31)
    sim, sc = make_pair(c, st, _lib.METHOD_ALLPAIRS)
    cur, _ = orc.step(c, sc, st)
    sim.step()
    c2 = orc.default_config(centering_factor=0.7, max_sight_angle=1.0, dt=0.004,
                            distance_weight_threshold=4.0)
    sim.set_config(py_config(c2))
    cur, _ = orc.step(c2, sc, cur)
    assert sim.step() == Duration.from_secs_f32(0.004)
    assert np.array_equal(bits(sim.read_state()), bits(cur))


def test_state_integrators(orc):
    """state.rs:166-185 and :218-280 through the C ABI, plus State<boid> Euler == inline Euler."""
    import ctypes as C
    lib = _lib.load()
    s = np.array([0, 0, 0, 0, 0, 1], f32)
    ds = np.array([0, 0, 1, 1, -1, 0], f32)
    out = np.zeros(6, f32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.fp_state_euler_combine(0, 6, P(s), P(ds), 0.5, P(out)) == 0
    assert list(out) == [0, 0, 0.5, 0.5, -0.5, 1.0]

    def deriv(v):  # y' = y - t^2 + 1 (state.rs:209-211)
        return np.array([v[0] - v[1] * v[1] + f32(1), 1, 0], f32)
    y = np.array([0.5, 0.0, 0.5], f32)
    h = f32(0.5)
    for golden in (1.425130208333333, 2.640859085770477, 4.009155464830968, 5.305471950534675):
        k1 = deriv(y)
        k2 = deriv(y + k1 * (h * f32(0.5)))
        k3 = deriv(y + k2 * (h * f32(0.5)))
        k4 = deriv(y + k3 * h)
        out = np.zeros(3, f32)
        assert lib.fp_state_rk4_combine(0, 3, P(y), P(k1), P(k2), P(k3), P(k4), float(h), P(out)) == 0
        assert np.array_equal(bits(out), bits(orc.state_rk4(y, 0.5, "examplefn")))
        assert abs(float(out[0]) - golden) < 0.005
        y = out
    # the flock as State<boid>
    c = orc.default_config()
    st = synth.uniform_flock(500, 30.0, seed=41)
    sim, sc = make_pair(c, st, _lib.METHOD_ALLPAIRS, TABLES)
    ref, _ = orc.step(c, sc, st)
    sim.state_euler(c.dt)
    assert np.array_equal(bits(sim.read_state()), bits(ref))
    # rk4 with frozen acceleration vs the oracle's generic rk4
    acc, _, _ = orc.accel_rows(c, sc, st)

    def dflock(v):
        v = v.reshape(-1, 6)
        return np.concatenate([v[:, 3:], acc], axis=1).reshape(-1)
    ref4 = orc.state_rk4(st.reshape(-1), 0.01, dflock).reshape(-1, 6)
    sim.write_state(st)
    sim.state_rk4(0.01)
    assert np.array_equal(bits(sim.read_state()), bits(ref4))


def test_instances_match_oracle(orc):
    st = synth.uniform_flock(1000, 30.0, seed=51)
    st[0, 3:] = [0, 0, 1]      # already along +z: identity
    st[1, 3:] = [0, 0, -2]     # anti-parallel: fallback axis
    st[2, 3:] = [3, 0, 0]
    sim, _ = make_pair(orc.default_config(), st, _lib.METHOD_ALLPAIRS)
    ref = orc.instances(st)
    got = sim.read_instances()
    assert np.array_equal(bits(got), bits(ref))
    assert list(got[0, 3:7]) == [1, 0, 0, 0] and got[0, 7] == f32(0.1)
    raw = sim.read_instances(raw=True)
    # model = T * R * S : column 3 is the position, |R col| = scale
    assert np.array_equal(raw[:, 12:15], st[:, :3]) and np.all(raw[:, 15] == 1)
    assert np.allclose(np.linalg.norm(raw[:, 0:3], axis=1), 0.1, atol=1e-6)
    inst = sim.get_boid_instances()
    assert len(inst) == 1000 and inst[5].scale == float(f32(0.1))


def test_instances_exported_into_mapped_and_device_buffers(orc):
    """fp_flock_export_instances: the GPU writes the Instance / InstanceRaw records straight into a
    buffer the caller maps -- pinned host memory (stores cross PCIe, no staging copy) and device
    memory (what a mapped graphics-interop vertex buffer is) -- same bits as the staged read-out;
    pageable memory is refused."""
    import torch
    st = synth.uniform_flock(5000, 60.0, seed=52)
    sim, _ = make_pair(orc.default_config(), st, _lib.METHOD_GRID)
    sim.step_many(3)
    for raw, width in ((False, 8), (True, 25)):
        ref = sim.read_instances(raw=raw)
        pinned = torch.zeros((len(st), width), dtype=torch.float32).pin_memory()
        sim.export_instances(pinned.data_ptr(), raw=raw)
        assert np.array_equal(bits(pinned.numpy()), bits(ref))
        dev = torch.zeros((len(st), width), dtype=torch.float32, device="cuda")
        sim.export_instances(dev.data_ptr(), raw=raw)
        assert np.array_equal(bits(dev.cpu().numpy()), bits(ref))
    pageable = np.zeros((len(st), 8), np.float32)
    with pytest.raises(_lib.FeriphysError):
        sim.export_instances(pageable.ctypes.data)


@pytest.mark.parametrize("n", [1, 37, 128, 129, 200, 256])
def test_small_kernels_bit_exact_vs_oracle(orc, n):
    """K4 in both forms -- lane-parallel (n <= 128: eight lanes evaluate a boid's pair terms, one
    thread adds them in index order) and one thread per boid (n <= 256) -- over 60 steps with
    tables, bit-identical to the reference loop."""
    c = orc.default_config()
    st = synth.uniform_flock(n, 14.0, seed=40 + n)
    sim, sc = make_pair(c, st, _lib.METHOD_SMALL, TABLES)
    assert sim.method_in_use() == _lib.METHOD_SMALL
    cur = st
    for _ in range(60):
        cur, _ = orc.step(c, sc, cur)
    sim.step_many(25)
    sim.step_many(35)
    assert np.array_equal(bits(sim.read_state()), bits(cur))
    c2 = orc.default_config(steering_overrides=1)
    sim.set_config(py_config(c2))
    sim.step_many(5)
    for _ in range(5):
        cur, _ = orc.step(c2, sc, cur)
    assert np.array_equal(bits(sim.read_state()), bits(cur))


def test_small_flock_write_step_read_through_mapped_memory(orc):
    """Demo-sized flocks: fp_flock_write_state leaves the rows in a pinned, device-mapped slot, the
    single-CTA step kernel ingests them itself and writes the advanced rows to mapped memory for
    fp_flock_read_state -- no copy, no conversion kernel.  Same bits as the oracle whichever way the
    calls interleave: write -> step -> read loops, a write that no step follows, taps between a write
    and a step, several steps per launch, a method change with rows still pending."""
    c = orc.default_config()
    st = synth.uniform_flock(110, 14.0, seed=61)
    for method in (_lib.METHOD_SMALL, _lib.METHOD_AUTO):
        sim, sc = make_pair(c, st, method, TABLES)
        cur = st.copy()
        for k in range(6):                       # the e2e loop of bench.py
            sim.write_state(cur)
            sim.step()
            cur, _ = orc.step(c, sc, cur)
            got = sim.read_state()
            assert np.array_equal(bits(got), bits(cur)), (method, k)
            assert np.array_equal(bits(sim.read_state()), bits(cur))      # a second read: same rows
        other = synth.uniform_flock(110, 14.0, seed=62)
        sim.write_state(other)                   # no step: the rows come back unchanged
        assert np.array_equal(bits(sim.read_state()), bits(other))
        sim.write_state(cur)
        ref_acc, _, _ = orc.accel_rows(c, sc, cur)
        assert np.array_equal(bits(sim.read_accel()), bits(ref_acc))      # a tap ingests pending rows
        sim.step_many(3)
        for _ in range(3):
            cur, _ = orc.step(c, sc, cur)
        assert np.array_equal(bits(sim.read_state()), bits(cur))
        sim.step()                               # a step straight after a read: state still on the device
        cur, _ = orc.step(c, sc, cur)
        assert np.array_equal(bits(sim.read_state()), bits(cur))
    # rows pending when the method changes: the all-pairs kernel must see them
    sim, sc = make_pair(c, st, _lib.METHOD_SMALL, TABLES)
    sim.write_state(other)
    sim.set_method(_lib.METHOD_ALLPAIRS)
    sim.step()
    ref, _ = orc.step(c, sc, other)
    assert np.array_equal(bits(sim.read_state()), bits(ref))
