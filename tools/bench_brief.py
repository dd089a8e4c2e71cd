#!/usr/bin/env python
"""Print the interesting fields of bench.py JSON lines read from stdin."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    out = {k: d.get(k) for k in ("impl", "value", "ms_per_step", "n_gpus", "gpu_launches")}
    out["workload"] = d["config"]["workload"][:2]
    for k in ("timing", "roofline", "e2e", "cpu_baseline", "clocks"):
        v = d.get(k)
        if isinstance(v, dict):
            out[k] = {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()
                      if a in ("value", "frac", "achieved", "sort_phase_ms_per_step", "influence_ms_per_step",
                               "cores", "sm_mhz", "reasons", "bound")}
    print(out)
