"""Deterministic synthetic flocks.

The reference seeds boids with an unseeded thread-local RNG
(``flocking.rs:63-95``: ``spawn + U[0,1)^3``, velocity ``U[0,1)^3``), so
"identical synthetic initial states" need an explicit generator.  This one is
``splitmix64`` keyed by ``(seed, boid index, component)`` with
``u01 = (x >> 40) * 2**-24`` -- the 24-bit shape rand 0.8's ``Standard``
distribution uses for ``f32``.  Everything is generated once on the host and
handed, bit for bit, to the oracle and to the GPU library alike.
"""
from __future__ import annotations

import numpy as np

SEED = 0xFE21F

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return z ^ (z >> np.uint64(31))


def u01(seed: int, n: int, components: int = 6, first: int = 0, index=None) -> np.ndarray:
    """``[n, components]`` float32 in [0,1), keyed by (seed, index, component).  ``index``: explicit
    boid indices instead of ``first .. first + n``."""
    rows = np.arange(first, first + n, dtype=np.uint64) if index is None else np.asarray(index, dtype=np.uint64)
    idx = rows[:, None] * np.uint64(components)
    comp = np.arange(components, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        key = _splitmix64(np.full((1, 1), seed, dtype=np.uint64)) ^ (idx + comp)
    x = _splitmix64(key)
    return ((x >> np.uint64(40)).astype(np.float32) * np.float32(2.0**-24)).astype(np.float32)


def uniform_flock(n: int, extent: float, seed: int = SEED, vel_lo: float = -1.0,
                  vel_hi: float = 1.0, first: int = 0, index=None) -> np.ndarray:
    """Positions U[0, extent)^3, velocities U[vel_lo, vel_hi)^3 -> ``[n, 6]`` float32
    (configs C2-C5).  ``first`` offsets the boid index so ranks can generate
    disjoint pieces of one global flock; ``index`` asks for the rows of given boids."""
    u = u01(seed, n, 6, first, index)
    n = len(u)
    s = np.empty((n, 6), dtype=np.float32)
    s[:, :3] = u[:, :3] * np.float32(extent)
    s[:, 3:] = np.float32(vel_lo) + u[:, 3:] * np.float32(vel_hi - vel_lo)
    return s


def spawn_flock(spawn_points, num_boids: int, seed: int = SEED) -> np.ndarray:
    """``Simulation::new`` (flocking.rs:73-85) with the jitter made explicit:
    ``num_boids / len(spawn_points)`` boids (integer division, remainder
    dropped) per spawn point, at ``spawn + U[0,1)^3`` with velocity ``U[0,1)^3``."""
    pts = np.asarray(spawn_points, dtype=np.float32).reshape(-1, 3)
    per = num_boids // len(pts)
    u = u01(seed, per * len(pts), 6)
    s = np.empty((per * len(pts), 6), dtype=np.float32)
    for k, p in enumerate(pts):
        blk = slice(k * per, (k + 1) * per)
        s[blk, :3] = p[None, :] + u[blk, :3]
        s[blk, 3:] = u[blk, 3:]
    return s


# --- the demo scene, demos/flocking.rs:92-156 (config C1) -------------------
DEMO_OBSTACLES = np.array([[-5.0, 0.0, 0.0, 4.0]], dtype=np.float32)  # ship at (-5,0,0), r = 1.0*4.0
DEMO_SIM1 = dict(spawn=[(25.0, 0.5, 0.0)], num_boids=110, lead_paths=(0,))
DEMO_SIM2 = dict(spawn=[(15.0, 10.0, 0.0), (25.0, 0.5, 0.0)], num_boids=110, lead_paths=(1, 2))


# --- config C5 tables ----------------------------------------
def c5_tables(extent: float = 1296.0, seed: int = SEED):
    """8 attractors (4 of mass +50, 4 of mass -50), 16 obstacle spheres r = 8,
    bbox = domain +-32.  Positions from the same keyed generator."""
    u = u01(seed ^ 0xA77, 24, 3)
    attractors = np.empty((8, 4), dtype=np.float32)
    attractors[:, :3] = u[:8] * np.float32(extent)
    attractors[:4, 3] = 50.0
    attractors[4:, 3] = -50.0
    obstacles = np.empty((16, 4), dtype=np.float32)
    obstacles[:, :3] = u[8:] * np.float32(extent)
    obstacles[:, 3] = 8.0
    bbox = np.array([-32.0, extent + 32.0] * 3, dtype=np.float32)
    return attractors, obstacles, bbox
