"""Helpers shared by the -m gpu parity tests: build the same scene for the
oracle and for the CUDA library (through the C ABI, via the Python mirror)."""
import numpy as np

from feriphys_b200 import _lib
from feriphys_b200.flocking import (BoundingBox, Config, Duration, Obstacle, PointAttractor,
                                    Simulation)
from oracle_lib import Scene

f32 = np.float32

TABLES = dict(
    attractors=np.array([[20, 20, 20, 5], [80, 10, 20, -4]], f32),
    obstacles=np.array([[40, 40, 40, 6], [10, 70, 30, 3]], f32),
    bbox=np.array([-50, 200, -50, 200, -50, 200], f32),
    leads=np.array([[5, 5, 5, 1, 0, 0, 10], [60, 9, 20, 0, 1, 0.5, 10]], f32),
)


class FixedLead:
    """A lead boid that never moves (for single-step parity of the lead term)."""

    def __init__(self, row):
        self._row = np.asarray(row, f32)

    def row(self):
        return self._row

    def step(self, dt):
        pass


def py_config(orc_cfg) -> Config:
    c = Config()
    for k in ("dt", "avoidance_factor", "centering_factor", "velocity_matching_factor",
              "distance_weight_threshold", "distance_weight_threshold_falloff", "max_sight_angle",
              "max_sight_angle_to_lead_boid"):
        setattr(c, k, getattr(orc_cfg, k))
    c.time_to_start_steering = Duration(orc_cfg.time_to_start_steering_secs,
                                        orc_cfg.time_to_start_steering_nanos)
    c.steering_overrides = bool(orc_cfg.steering_overrides)
    return c


def make_pair(orc_cfg, state, method, tables=None, device=0, numerics=_lib.NUMERICS_EXACT):
    """-> (Simulation on the GPU, oracle Scene) describing the same flock.  EXACT numerics unless
    asked otherwise: most parity tests compare bit patterns."""
    t = tables or {}
    sim = Simulation.from_state(
        state,
        bounding_box=(BoundingBox(t["bbox"][0:2], t["bbox"][2:4], t["bbox"][4:6])
                      if "bbox" in t else None),
        lead_boids=[FixedLead(r) for r in t["leads"]] if "leads" in t else None,
        obstacles=[Obstacle(o[:3], float(o[3])) for o in t["obstacles"]] if "obstacles" in t else None,
        attractors=([PointAttractor(a[:3], float(a[3])) for a in t["attractors"]]
                    if "attractors" in t else None),
        method=method, device=device, numerics=numerics)
    sim.set_config(py_config(orc_cfg))
    scene = Scene(leads=t.get("leads"), attractors=t.get("attractors"),
                  obstacles=t.get("obstacles"), bbox=t.get("bbox"))
    return sim, scene


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def rel_err(got, ref, floor=1e-3):
    """max over boids of |got - ref| / max(|ref|, floor), vector norms."""
    num = np.linalg.norm(got.astype(np.float64) - ref.astype(np.float64), axis=-1)
    den = np.maximum(np.linalg.norm(ref.astype(np.float64), axis=-1), floor)
    return float((num / den).max())
