#!/usr/bin/env python
"""GPU: step a grid flock and print one JSON line with a hash of the final state.

Used to compare the walk on standing candidate lists (fp_walk_nl.cu, the default) with the plain
staged walk (FP_NL=0), which must agree bit for bit under EXACT numerics as long as both keep the
same binnings (pin the skin with FP_SKIN): run it once per setting and compare `sha256`.

    python tools/nl_state_hash.py [n] [extent] [steps] [seed] [blob]

`blob` > 0 adds that many boids inside a ball of radius 6 (thousands of neighbours each: the
candidate lists overflow and the library must fall back to the staged walk)."""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

from feriphys_b200 import _lib, synth  # noqa: E402
from feriphys_b200.flocking import Simulation  # noqa: E402


def main():
    a = sys.argv[1:]
    n = int(a[0]) if len(a) > 0 else 200_000
    extent = float(a[1]) if len(a) > 1 else 470.0      # ~33 neighbours within reach, as C3 / C4
    steps = int(a[2]) if len(a) > 2 else 120
    seed = int(a[3]) if len(a) > 3 else 7
    blob = int(a[4]) if len(a) > 4 else 0
    st = synth.uniform_flock(n, extent, seed=seed)
    if blob:
        rng = np.random.default_rng(seed)
        b = synth.uniform_flock(blob, 1.0, seed=seed + 1)
        d = rng.normal(size=(blob, 3))
        d *= (6.0 * rng.random(blob) ** (1 / 3) / np.linalg.norm(d, axis=1))[:, None]
        b[:, :3] = (extent / 2 + d).astype(np.float32)
        st = np.concatenate([st, b]).astype(np.float32)
    sim = Simulation.from_state(st, method=_lib.METHOD_GRID)
    sim.step_many(3)
    sim.sync()
    t0 = time.perf_counter()
    chunk = int(os.environ.get("NL_CHUNK", "0"))     # diagnostic: sync every `chunk` steps, report progress
    if chunk:
        done = 0
        while done < steps:
            k = min(chunk, steps - done)
            try:
                sim.step_many(k)
                sim.sync()
            except Exception as e:
                print(f"failed within steps {done + 1}..{done + k} (after the 3 warm-up steps): {e}", file=sys.stderr)
                raise
            done += k
            print("progress", done, "rebin_info", sim.rebin_info(), "grid", sim.grid_info(), file=sys.stderr)
    else:
        sim.step_many(steps)
        sim.sync()
    ms = 1e3 * (time.perf_counter() - t0) / steps
    out = sim.read_state()
    skin, nsteps, rebins, replayed = sim.rebin_info()
    print(json.dumps({"lists": os.environ.get("FP_NL", "1") != "0", "boids": len(st), "steps": steps,
                      "sha256": hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest(),
                      "skin": skin, "rebins": int(rebins), "replayed": int(replayed), "wall_ms_per_step": ms,
                      "finite": bool(np.isfinite(out).all())}))


if __name__ == "__main__":
    main()
