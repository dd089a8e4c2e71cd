#!/usr/bin/env python
"""Diagnostic (GPU): step a standing binning one step at a time against the oracle and
describe the first step whose bits differ.  python tools/diag_rebin.py [skin] [max_steps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

from feriphys_b200 import _lib, synth  # noqa: E402
from gpu_util import bits, make_pair  # noqa: E402
from oracle_lib import oracle  # noqa: E402


def main():
    skin = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    nmax = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    orc = oracle()
    nt = os.cpu_count() or 1
    c = orc.default_config()
    st = synth.uniform_flock(20000, 200.0, seed=91)
    sim, sc = make_pair(c, st, _lib.METHOD_GRID)
    sim.set_rebin(skin=skin)
    sim.read_neighbors()
    idx, cur = sim.read_local()
    idx = idx.astype(np.int64)
    binned = cur.copy()
    dims, cell, _ = sim.grid_info()
    origin = st[:, :3].min(axis=0)
    print("variant", os.environ.get("FP_WALK_VARIANT"), "skin", skin, "dims", dims, "cell", cell, flush=True)
    for step in range(1, nmax + 1):
        prev = cur
        sim.step()
        cur, _ = orc.step(c, sc, prev, threads=nt, grid=True)
        idx_k, got = sim.read_local()
        info = sim.rebin_info()
        if not np.array_equal(idx_k.astype(np.int64), idx):
            print("step", step, "listing changed; rebin_info", info)
            return
        bad = np.nonzero((bits(got) != bits(cur)).any(axis=1))[0]
        if len(bad):
            print("step", step, "rows differing:", len(bad), "rebin_info", info)
            cb = np.clip(np.floor((binned[:, :3] - origin) / np.float32(cell)).astype(int), 0, np.array(dims) - 1)
            for i in bad[:8]:
                d = prev[:, :3] - prev[i, :3]
                m2 = (d * d).sum(axis=1)
                nb = np.nonzero(m2 < 16.0 * 16.0)[0]
                nb = nb[nb != i]
                far = [j for j in nb if np.abs(cb[j] - cb[i]).max() > 1]
                print(" slot", i, "dv", (got[i, 3:] - cur[i, 3:]), "dp", (got[i, :3] - cur[i, :3]),
                      "moved", np.linalg.norm(prev[i, :3] - binned[i, :3]), "neighbours", len(nb),
                      "outside home 27 cells:", [(int(j), float(np.sqrt(m2[j])),
                                                  float(np.linalg.norm(prev[j, :3] - binned[j, :3]))) for j in far])
            return
    print("no difference in", nmax, "steps; rebin_info", sim.rebin_info())


if __name__ == "__main__":
    main()
