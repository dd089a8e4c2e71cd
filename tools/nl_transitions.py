#!/usr/bin/env python
"""GPU: one flock taken through the state transitions that surround the walk -- taps that re-bin,
an outrun plan with replays, a config switch (steering overrides on and off), tables, a new state
from the host, a detour through the all-pairs kernel -- printing a hash of the state after each
stage.  Run once with the candidate lists (default) and once without (FP_NL=0), the skin pinned by
FP_SKIN: both must print the same lines."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

from feriphys_b200 import _lib, synth  # noqa: E402
from feriphys_b200.flocking import Config, Obstacle, PointAttractor, Simulation  # noqa: E402

f32 = np.float32


def main():
    v = "lists" if os.environ.get("FP_NL", "1") != "0" else "staged walk"
    stage = [0]

    def show(sim, what):
        st = sim.read_state()
        stage[0] += 1
        print(f"{stage[0]:2d} {what:34s} {hashlib.sha256(st.tobytes()).hexdigest()[:16]} "
              f"rebin_info {sim.rebin_info()[1:]} finite {bool(np.isfinite(st).all())}", flush=True)

    print("walk:", v)
    st = synth.uniform_flock(40000, 270.0, seed=21)
    sim = Simulation.from_state(st, method=_lib.METHOD_GRID,
                                attractors=[PointAttractor(np.array([-60, 80, 80], f32), 2.0e4)],
                                obstacles=[Obstacle(np.array([100, 100, 100], f32), 9.0)])
    sim.step_many(40);                       show(sim, "40 steps, attractor + obstacle")
    sim.read_neighbors(); sim.step_many(25); show(sim, "tap, then 25 steps")
    sim.read_accel(); sim.step(); sim.step(); show(sim, "accel tap, 2 single steps")
    sim.set_rebin(skin=0.1, plan_scale=50.0)
    sim.step_many(120);                      show(sim, "outrun plan (replays), 120 steps")
    sim.set_rebin()
    c = Config(); c.steering_overrides = True
    sim.set_config(c); sim.step_many(10);    show(sim, "steering overrides on, 10 steps")
    c.steering_overrides = False; c.max_sight_angle = 2.0
    sim.set_config(c); sim.step_many(30);    show(sim, "overrides off, FOV 2.0, 30 steps")
    sim.write_state(synth.uniform_flock(40000, 270.0, seed=22))
    sim.step_many(30);                       show(sim, "new state from the host, 30 steps")
    sim.set_method(_lib.METHOD_ALLPAIRS); sim.step()
    sim.set_method(_lib.METHOD_GRID); sim.step_many(30); show(sim, "all-pairs detour, 30 grid steps")
    sim.state_euler(1e-3); sim.step_many(20); show(sim, "State::euler_step, 20 steps")
    sim.step_many(300);                      show(sim, "300 steps (crosses a re-fit)")


if __name__ == "__main__":
    main()
