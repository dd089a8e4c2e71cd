#!/usr/bin/env python
"""Regenerates tests/golden/flocking_golden.npz.

The reference (jalberse/feriphys) is Rust and cannot be run in this image, and it holds
no golden vectors of its own for the flocking path, so these fixtures are
frozen outputs of the CPU oracle (oracle/flock_oracle.c) on seeded inputs.  They pin the
oracle against accidental change (tests/test_golden.py, CPU) and give the CUDA kernels a
second, immutable target (same file, -m gpu).  Run from the repository root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from feriphys_b200 import synth  # noqa: E402
from oracle_lib import Scene, oracle  # noqa: E402

f32 = np.float32
TABLES = dict(
    attractors=np.array([[20, 20, 20, 5], [80, 10, 20, -4]], f32),
    obstacles=np.array([[40, 40, 40, 6], [10, 70, 30, 3]], f32),
    bbox=np.array([-50, 200, -50, 200, -50, 200], f32),
    leads=np.array([[5, 5, 5, 1, 0, 0, 10], [60, 9, 20, 0, 1, 0.5, 10]], f32),
)


def main():
    orc = oracle()
    out = {}
    # 1. the demo scene, sim 1 and sim 2 (demos/flocking.rs:92-156): state after 200 steps
    for which, scene in ((1, synth.DEMO_SIM1), (2, synth.DEMO_SIM2)):
        kinds = list(scene["lead_paths"])
        st = synth.spawn_flock(scene["spawn"], scene["num_boids"])
        cfg = orc.default_config()
        leads, times = orc.make_leads(kinds)
        cur = st
        for _ in range(200):
            cur, _ = orc.step(cfg, Scene(leads=leads, obstacles=synth.DEMO_OBSTACLES), cur)
            orc.step_leads(leads, times, kinds, cfg.dt)
        out[f"demo{which}_state0"] = st
        out[f"demo{which}_state200"] = cur
        out[f"demo{which}_leads200"] = leads
    # 2. a 600-boid flock with every table: accelerations by component, neighbour sets, 10 steps
    st = synth.uniform_flock(600, 45.0, seed=1234)
    cfg = orc.default_config()
    sc = Scene(**TABLES)
    total, comp, flags = orc.accel_rows(cfg, sc, st)
    cnt, hsh, _ = orc.neighbors_rows(cfg, st)
    cur = st
    for _ in range(10):
        cur, _ = orc.step(cfg, sc, cur)
    out.update(flock_state0=st, flock_accel=total, flock_comp=comp, flock_flags=flags, flock_ncount=cnt,
               flock_nhash=hsh, flock_state10=cur)
    # 3. a narrow-FOV, short-range configuration
    cfg2 = orc.default_config(max_sight_angle=0.8, distance_weight_threshold=5.0,
                              distance_weight_threshold_falloff=2.5, centering_factor=0.3, dt=0.002)
    st2 = synth.uniform_flock(500, 30.0, seed=4321)
    t2, _, _ = orc.accel_rows(cfg2, None, st2)
    c2, h2, _ = orc.neighbors_rows(cfg2, st2)
    out.update(narrow_state0=st2, narrow_accel=t2, narrow_ncount=c2, narrow_nhash=h2)
    path = os.path.join(ROOT, "tests", "golden", "flocking_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
