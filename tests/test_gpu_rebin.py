"""-m gpu tests of the grid path's lazy re-binning (fp_flock_set_rebin / fp_flock_rebin_info).

One binning (sort by cell) serves many steps: the cell edge carries a skin, the walk takes
each boid's home cell from the key it was binned under, and the device bounds every boid's
displacement since the binning with max|v| * dt.  What must hold:
  * while a binning stands, the library's internal listing does not change and every step is
    BIT-IDENTICAL to Simulation::step (flocking.rs:97-131) run by the oracle on that listing --
    a single missed or extra neighbour would change bits, so this also pins the neighbour sets
    of every intermediate step, including boids that have drifted out of their home cell;
  * a plan that is too optimistic is caught on the device, the voided steps are replayed after
    a fresh binning, and the result does not depend on the plan beyond f32 summation order.
"""
import os

import numpy as np
import pytest

from feriphys_b200 import _lib, synth
from gpu_util import bits, make_pair

pytestmark = pytest.mark.gpu
f32 = np.float32
NT = os.cpu_count() or 1


def _listing(sim):
    idx, internal = sim.read_local()
    return idx.astype(np.int64), internal


def test_standing_binning_is_bit_identical_to_the_reference_loop(orc):
    c = orc.default_config()
    st = synth.uniform_flock(20000, 200.0, seed=91)
    sim, sc = make_pair(c, st, _lib.METHOD_GRID)   # plain flock: speeds stay near |v| <= 1.73
    sim.set_rebin(skin=1.0)          # fixed skin: ~200 steps per binning at |v| <= 1.73, dt = 1e-3
    sim.read_neighbors()             # a tap bins the flock where it stands
    idx, cur = _listing(sim)
    skin, steps0, rebins0, _ = sim.rebin_info()
    assert skin == 1.0
    _, cell, _ = sim.grid_info()
    assert cell > 17.0               # reach 16 (+1/512) + skin
    total = 0
    for k in (1, 9, 60, 80):
        sim.step_many(k)
        for _ in range(k):
            cur, _ = orc.step(c, sc, cur, threads=NT, grid=True)
        total += k
        idx_k, got = _listing(sim)
        assert np.array_equal(idx_k, idx), "the listing moved although the binning stands"
        assert np.array_equal(bits(got), bits(cur)), f"step {total} differs from the reference loop"
    _, steps1, rebins1, replayed = sim.rebin_info()
    assert steps1 - steps0 == total and rebins1 == rebins0 and replayed == 0
    # by now a good share of the boids sit outside the cell they were binned under
    moved = np.abs(cur[:, :3] - st[idx][:, :3]).max()
    assert moved > 0.1
    # a fresh binning of the evolved state sees exactly the oracle's neighbour sets
    state = sim.read_state()
    rc, rh, _ = orc.neighbors_rows(c, state, threads=NT, grid=True)
    gc, gh = sim.read_neighbors()
    assert np.array_equal(gc, rc) and np.array_equal(gh, rh)


def test_automatic_skin_rebins_on_schedule(orc):
    c = orc.default_config()
    st = synth.uniform_flock(30000, 240.0, seed=92)
    sim, sc = make_pair(c, st, _lib.METHOD_GRID)
    sim.step_many(300)             # (crosses the re-fit of the grid every 256 steps)
    skin, steps, rebins, replayed = sim.rebin_info()
    assert steps == 300 and replayed == 0
    assert 0.02 < skin < 2.0
    assert 2 <= rebins <= 90, f"{rebins} binnings for 300 steps"
    cur = st
    for _ in range(300):
        cur, _ = orc.step(c, sc, cur, threads=NT, grid=True)
    got = sim.read_state()
    scale = np.maximum(1.0, np.linalg.norm(cur[:, :3], axis=1))
    assert (np.linalg.norm(got[:, :3] - cur[:, :3], axis=1) / scale).max() <= 1e-4
    # skin 0 is the classic scheme: one binning per step, same trajectory up to summation order
    ref, _ = make_pair(c, st, _lib.METHOD_GRID)
    ref.set_rebin(skin=0.0)
    ref.step_many(300)
    assert ref.rebin_info()[2] == 300
    other = ref.read_state()
    assert (np.linalg.norm(got[:, :3] - other[:, :3], axis=1) / scale).max() <= 1e-4


def test_outrun_plan_is_voided_on_the_device_and_replayed(orc):
    # a heavy attractor beside the flock speeds the flock up step after step; the host plans 50x too
    # many steps per binning, so only the device-side bound keeps the neighbour sets exact
    c = orc.default_config()
    st = synth.uniform_flock(12000, 160.0, seed=93)
    tables = dict(attractors=np.array([[-60, 80, 80, 2.0e4]], f32))
    sim, sc = make_pair(c, st, _lib.METHOD_GRID, tables)
    sim.set_rebin(skin=0.1, plan_scale=50.0)   # a small skin: the speed-up outruns it within the run
    n = 120
    sim.step_many(n)
    skin, steps, rebins, replayed = sim.rebin_info()
    assert steps == n
    assert replayed > 0 and rebins >= 2
    cur = st
    for _ in range(n):
        cur, _ = orc.step(c, sc, cur, threads=NT, grid=True)
    got = sim.read_state()
    assert np.abs(cur[:, 3:]).max() > 3.0      # the flock did get faster
    scale = np.maximum(1.0, np.linalg.norm(cur[:, :3], axis=1))
    assert (np.linalg.norm(got[:, :3] - cur[:, :3], axis=1) / scale).max() <= 1e-4
    rc, rh, _ = orc.neighbors_rows(c, got, threads=NT, grid=True)
    gc, gh = sim.read_neighbors()
    assert np.array_equal(gc, rc) and np.array_equal(gh, rh)
    # stepping one at a time with reads in between takes the same decisions in another rhythm
    one, _ = make_pair(c, st, _lib.METHOD_GRID, tables)
    one.set_rebin(skin=0.1, plan_scale=50.0)
    for _ in range(n):
        one.step()
        one.sync()
    again = one.read_state()
    assert (np.linalg.norm(again[:, :3] - cur[:, :3], axis=1) / scale).max() <= 1e-4


def test_state_changes_invalidate_the_binning(orc):
    c = orc.default_config()
    st = synth.uniform_flock(9000, 130.0, seed=94)
    sim, sc = make_pair(c, st, _lib.METHOD_GRID)
    sim.set_rebin(skin=0.5)
    sim.step_many(3)
    r0 = sim.rebin_info()[2]
    # a new state from the host must be binned afresh ...
    st2 = synth.uniform_flock(9000, 130.0, seed=95)
    sim.write_state(st2)
    sim.step()
    assert sim.rebin_info()[2] == r0 + 1
    one, _ = orc.step(c, sc, st2, threads=NT, grid=True)
    got = sim.read_state()
    assert np.abs(got - one).max() <= 1e-5 * max(1.0, float(np.abs(one).max()))
    # ... and so must one that went through the all-pairs kernel (caller order) in between
    sim.set_method(_lib.METHOD_ALLPAIRS)
    sim.step()
    sim.set_method(_lib.METHOD_GRID)
    sim.step()
    assert sim.rebin_info()[2] == r0 + 2
    cur = one
    for _ in range(2):
        cur, _ = orc.step(c, sc, cur, threads=NT, grid=True)
    got = sim.read_state()
    assert np.abs(got - cur).max() <= 1e-5 * max(1.0, float(np.abs(cur).max()))
