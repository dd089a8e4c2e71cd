"""Host-side mirror of feriphys's flocking API over the CUDA library.

Same names, argument meaning and error behaviour as the reference
(``src/simulation/flocking/{flocking,boid,obstacle}.rs``, ``parametric.rs``,
``point_attractor.rs``, ``bounding_box.rs``), so a caller of
``flocking::Simulation`` finds ``Simulation.new / step / get_timestep /
sync_sim_config_from_ui / get_boid_instances`` unchanged.  Everything numeric
happens in ``libferiphys_cuda.so``; this module only owns what cannot cross a
C ABI -- the lead boids' ``fn(f32) -> Vector3`` paths, which are evaluated here
in the reference's order (one-step lag, f32 time accumulation) and
uploaded as tables.

ADDITIONS over the reference (its state is neither injectable nor readable,
SURVEY F3/F4): ``Simulation.from_state``, ``read_state``, ``step_many`` and the
debug taps ``read_accel`` / ``read_neighbors`` / ``pair_census``.
"""
from __future__ import annotations

import ctypes as C
import ctypes.util
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np

from . import _lib, synth
from ._lib import FpConfig, check, f32c, ptr

f32 = np.float32
PI = f32(3.14159274101257324)  # std::f32::consts::PI

_libm = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n in ("cosf", "sinf"):
    getattr(_libm, _n).restype = C.c_float
    getattr(_libm, _n).argtypes = [C.c_float]


def cosf(x) -> np.float32:
    """``f32::cos`` -- the platform libm's cosf, as Rust calls it."""
    return f32(_libm.cosf(float(f32(x))))


def sinf(x) -> np.float32:
    return f32(_libm.sinf(float(f32(x))))


class Panic(RuntimeError):
    """Raised where the Rust reference would ``panic!``."""


@dataclass(frozen=True, order=True)
class Duration:
    """``std::time::Duration``: whole seconds + nanoseconds."""
    secs: int = 0
    nanos: int = 0

    @staticmethod
    def from_secs(s: int) -> "Duration":
        return Duration(int(s), 0)

    @staticmethod
    def from_millis(ms: int) -> "Duration":
        return Duration(ms // 1000, (ms % 1000) * 1_000_000)

    @staticmethod
    def from_micros(us: int) -> "Duration":
        return Duration(us // 1_000_000, (us % 1_000_000) * 1000)

    @staticmethod
    def from_secs_f32(x) -> "Duration":
        """Exact value times 1e9 rounded to the nearest ns, ties to even; panics on
        negative, NaN and overflow like ``Duration::from_secs_f32``."""
        x = float(f32(x))
        if x < 0.0:
            raise Panic("can not convert float seconds to Duration: value is negative")
        if not x < 18446744073709551616.0:
            raise Panic("can not convert float seconds to Duration: value is either too big or NaN")
        if x >= 8388608.0:
            return Duration(int(x), 0)
        ns = int(round(x * 1e9))  # product exact in binary64; round() is half-to-even
        return Duration(ns // 1_000_000_000, ns % 1_000_000_000)

    def as_secs_f32(self) -> np.float32:
        return f32(f32(self.secs) + f32(self.nanos) / f32(1e9))

    def is_zero(self) -> bool:
        return self.secs == 0 and self.nanos == 0


@dataclass
class Config:
    """``flocking::Config`` with ``Default`` values (flocking.rs:15-51)."""
    dt: float = float(Duration.from_millis(1).as_secs_f32())
    avoidance_factor: float = 1.0
    centering_factor: float = float(f32(0.1))
    velocity_matching_factor: float = 0.5
    distance_weight_threshold: float = 15.0
    distance_weight_threshold_falloff: float = 1.0
    max_sight_angle: float = float(PI / f32(2.0))
    max_sight_angle_to_lead_boid: float = float(PI)
    time_to_start_steering: Duration = Duration.from_secs(4)
    steering_overrides: bool = False

    def to_c(self) -> FpConfig:
        c = FpConfig()
        for k in ("dt", "avoidance_factor", "centering_factor", "velocity_matching_factor",
                  "distance_weight_threshold", "distance_weight_threshold_falloff",
                  "max_sight_angle", "max_sight_angle_to_lead_boid"):
            setattr(c, k, float(f32(getattr(self, k))))
        c.time_to_start_steering_secs = self.time_to_start_steering.secs
        c.time_to_start_steering_nanos = self.time_to_start_steering.nanos
        c.steering_overrides = 1 if self.steering_overrides else 0
        return c


class Parametric:
    """``parametric.rs``: a curve in R3; ``step`` returns ``path(t)`` THEN advances t."""

    def __init__(self, path: Callable[[np.float32], Sequence[float]]):
        self.path = path
        self.curr_time = f32(0.0)

    def step(self, dt) -> np.ndarray:
        position = np.asarray(self.path(self.curr_time), dtype=np.float32)
        self.curr_time = f32(self.curr_time + f32(dt))
        return position


class LeadBoid:
    """``boid.rs:13-53``: follows a parametric path; weight 10."""

    def __init__(self, path: Callable[[np.float32], Sequence[float]]):
        self.parametric = Parametric(path)
        self._position = np.asarray(path(f32(0.0)), dtype=np.float32)
        self._velocity = np.zeros(3, np.float32)
        self._weight = f32(10.0)

    new = classmethod(lambda cls, path: cls(path))

    def position(self):
        return self._position

    def velocity(self):
        return self._velocity

    def weight(self):
        return self._weight

    def step(self, dt: Duration) -> None:
        if dt.is_zero():
            return
        s = dt.as_secs_f32()
        new_position = self.parametric.step(s)
        self._velocity = ((new_position - self._position) / s).astype(np.float32)
        self._position = new_position

    def row(self) -> np.ndarray:
        return np.concatenate([self._position, self._velocity, [self._weight]]).astype(np.float32)


@dataclass
class Obstacle:
    """``obstacle.rs:11-14``: bounding sphere."""
    position: Sequence[float]
    radius: float

    @staticmethod
    def from_entity(instances, radius: float):
        """``Obstacle::from_entity`` (obstacle.rs:48-59): one sphere per instance,
        ``radius = instance.scale * radius``."""
        return [Obstacle(i.position, float(f32(i.scale) * f32(radius))) for i in instances]


@dataclass
class PointAttractor:
    """``point_attractor.rs:9-12``; a negative mass repels."""
    position: Sequence[float]
    mass: float


@dataclass
class BoundingBox:
    """``bounding_box.rs:5-9``: three ``Range<f32>`` as (start, end) pairs."""
    x_range: Sequence[float]
    y_range: Sequence[float]
    z_range: Sequence[float]


@dataclass
class Instance:
    """``graphics/instance.rs:7-11``; rotation is a quaternion (s, x, y, z)."""
    position: np.ndarray
    rotation: np.ndarray
    scale: float = 1.0


class Simulation:
    """``flocking::Simulation`` (flocking.rs:53-246) on one B200."""

    def __init__(self, initial_positions, num_boids: int, bounding_box: Optional[BoundingBox] = None,
                 lead_boids: Optional[list] = None, obstacles: Optional[list] = None,
                 attractors: Optional[list] = None, *, seed: int = synth.SEED, device: int = 0,
                 method: int = _lib.METHOD_AUTO, numerics: Optional[int] = None,
                 _state: Optional[np.ndarray] = None):
        """``Simulation::new`` (flocking.rs:63-95).  The reference jitters spawn points
        with an unseeded RNG; here the jitter comes from ``synth.spawn_flock(seed)``."""
        if _state is None:
            if len(initial_positions) == 0:
                raise Panic("attempt to divide by zero")  # flocking.rs:74-76
            _state = synth.spawn_flock(initial_positions, num_boids, seed)
        self._lib = _lib.load()
        self.config = Config()
        self.lead_boids = lead_boids
        self.bounding_box = bounding_box
        self.obstacles = obstacles
        self.attractors = attractors
        state = f32c(_state, (-1, 6))
        self._h = self._create(state, device)
        check(self._lib.fp_flock_set_method(self._h, method))
        if numerics is not None:
            check(self._lib.fp_flock_set_numerics(self._h, numerics))
        self._push_tables()

    def _create(self, state: np.ndarray, device: int):
        """-> library handle; sets ``self._n`` (rows of the per-boid outputs)."""
        self._n = len(state)
        h = C.c_void_p()
        cfg = self.config.to_c()
        check(self._lib.fp_flock_create(C.byref(h), C.byref(cfg), self._n, ptr(state), device))
        return h

    new = classmethod(lambda cls, *a, **k: cls(*a, **k))

    @classmethod
    def from_state(cls, state, bounding_box=None, lead_boids=None, obstacles=None, attractors=None,
                   **kw) -> "Simulation":
        """ADDITION (F3): explicit ``[n, 6]`` initial state (px py pz vx vy vz)."""
        return cls(None, 0, bounding_box, lead_boids, obstacles, attractors, _state=state, **kw)

    # ---- tables ---------------------------------------------------------------
    def _push_tables(self) -> None:
        L = self._lib
        if self.bounding_box is not None:
            b = self.bounding_box
            arr = f32c([b.x_range[0], b.x_range[1], b.y_range[0], b.y_range[1], b.z_range[0],
                        b.z_range[1]])
            check(L.fp_flock_set_bbox(self._h, ptr(arr)))
        else:
            check(L.fp_flock_set_bbox(self._h, None))
        att = f32c([[*a.position, a.mass] for a in (self.attractors or [])], (-1, 4))
        check(L.fp_flock_set_attractors(self._h, len(att), ptr(att) if len(att) else None))
        obs = f32c([[*o.position, o.radius] for o in (self.obstacles or [])], (-1, 4))
        check(L.fp_flock_set_obstacles(self._h, len(obs), ptr(obs) if len(obs) else None))
        self._push_leads()

    def _lead_rows(self) -> np.ndarray:
        return f32c([l.row() for l in (self.lead_boids or [])], (-1, 7))

    def _push_leads(self) -> None:
        rows = self._lead_rows()
        check(self._lib.fp_flock_set_leads(self._h, len(rows), ptr(rows) if len(rows) else None))

    # ---- reference API ----------------------------------------------------------
    def step(self) -> Duration:
        """``Simulation::step`` (flocking.rs:97-131)."""
        if self.lead_boids:
            self._push_leads()
        check(self._lib.fp_flock_step(self._h, 1))
        dt = self.get_timestep()
        if self.lead_boids:
            for lead in self.lead_boids:
                lead.step(dt)
        return dt

    def step_many(self, nsteps: int) -> Duration:
        """ADDITION: ``nsteps`` steps in one library call; the lead boids' rows are
        tabulated up front exactly as ``nsteps`` successive ``step()`` calls would."""
        if nsteps <= 0:
            return Duration()
        if self.lead_boids:
            dt = Duration.from_secs_f32(self.config.dt)
            table = np.empty((nsteps, len(self.lead_boids), 7), np.float32)
            for s in range(nsteps):
                table[s] = self._lead_rows()
                for lead in self.lead_boids:
                    lead.step(dt)
            check(self._lib.fp_flock_set_lead_table(self._h, nsteps, table.shape[1], ptr(table)))
        check(self._lib.fp_flock_step(self._h, nsteps))
        if self.lead_boids:
            self._push_leads()  # drops the table, leaves the current rows in force
        return self.get_timestep()

    def get_timestep(self) -> Duration:
        return Duration.from_secs_f32(self.config.dt)

    def sync_sim_config_from_ui(self, ui) -> None:
        """flocking.rs:215-228: ``ui.get_gui_state_mut()`` returns a ``Config``."""
        self.set_config(ui.get_gui_state_mut())

    def set_config(self, cfg: Config) -> None:
        self.config = Config(**{k: getattr(cfg, k) for k in Config.__dataclass_fields__})
        c = self.config.to_c()
        check(self._lib.fp_flock_set_config(self._h, C.byref(c)))

    def get_boid_instances(self):
        """flocking.rs:230-245 -> list of :class:`Instance` (scale 0.1)."""
        raw = self.read_instances()
        return [Instance(r[:3].copy(), r[3:7].copy(), float(r[7])) for r in raw]

    # ---- additions ----------------------------------------------------------------
    def __len__(self) -> int:
        return self._n

    def set_method(self, method: int) -> None:
        check(self._lib.fp_flock_set_method(self._h, method))

    def method_in_use(self) -> int:
        m = C.c_int(0)
        check(self._lib.fp_flock_get_method(self._h, C.byref(m)))
        return m.value

    def set_numerics(self, numerics: int) -> None:
        """``_lib.NUMERICS_EXACT`` (bit-identical arithmetic) or ``_lib.NUMERICS_FAST`` (neighbour
        sets bit-exact, forces fused: accelerations within ~1e-6 relative)."""
        check(self._lib.fp_flock_set_numerics(self._h, numerics))

    def numerics(self):
        """-> (requested, in use)"""
        a, b = C.c_int(0), C.c_int(0)
        check(self._lib.fp_flock_get_numerics(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def sync(self) -> None:
        check(self._lib.fp_flock_sync(self._h))

    def read_state(self) -> np.ndarray:
        out = np.empty((self._n, 6), np.float32)
        check(self._lib.fp_flock_read_state(self._h, ptr(out)))
        return out

    def read_local(self):
        """-> (index, state) in the library's INTERNAL order (cell-sorted after a grid
        step): ``state[k]`` is the boid the caller knows as ``index[k]``."""
        n = C.c_uint64(0)
        check(self._lib.fp_flock_local_len(self._h, C.byref(n)))
        idx = np.empty(n.value, np.uint64)
        st = np.empty((n.value, 6), np.float32)
        check(self._lib.fp_flock_read_local(self._h, ptr(idx), ptr(st)))
        return idx, st

    def write_local(self, index, state) -> None:
        """New values for the rows ``read_local`` listed, in that order (works on sharded flocks)."""
        idx = np.ascontiguousarray(index, np.uint64)
        st = f32c(state, (len(idx), 6))
        check(self._lib.fp_flock_write_local(self._h, len(idx), ptr(idx), ptr(st)))

    def write_state(self, state) -> None:
        st = f32c(state, (self._n, 6))
        check(self._lib.fp_flock_write_state(self._h, ptr(st)))

    def read_instances(self, raw: bool = False) -> np.ndarray:
        out = np.empty((self._n, 25 if raw else 8), np.float32)
        fn = self._lib.fp_flock_read_instances_raw if raw else self._lib.fp_flock_read_instances
        check(fn(self._h, ptr(out)))
        return out

    def export_instances(self, dst_ptr: int, raw: bool = False) -> None:
        """``get_boid_instances`` written by the GPU into memory the caller maps (device memory or
        pinned host memory): ``dst_ptr`` is the address, ``n x 8`` (or ``n x 25``) floats."""
        check(self._lib.fp_flock_export_instances(self._h, C.c_void_p(dst_ptr), 1 if raw else 0))

    def read_accel(self, components: bool = False):
        """ADDITION (debug tap): the acceleration the NEXT step would apply -- with the lead boids
        where they stand now (``step()`` leaves the library holding the rows of the step before)."""
        if self.lead_boids:
            self._push_leads()
        acc = np.empty((self._n, 3), np.float32)
        comp = np.empty((self._n, 5, 3), np.float32) if components else None
        check(self._lib.fp_flock_read_accel(self._h, ptr(acc), ptr(comp)))
        return (acc, comp) if components else acc

    def read_neighbors(self):
        cnt = np.empty(self._n, np.uint32)
        hsh = np.empty(self._n, np.uint64)
        check(self._lib.fp_flock_read_neighbors(self._h, ptr(cnt), ptr(hsh)))
        return cnt, hsh

    def pair_census(self) -> np.ndarray:
        out = np.zeros(4, np.uint64)
        check(self._lib.fp_flock_pair_census(self._h, ptr(out)))
        return out

    def status(self) -> int:
        v = C.c_uint32(0)
        check(self._lib.fp_flock_status(self._h, C.byref(v)))
        return v.value

    def set_grid_domain(self, lo, hi) -> None:
        lo, hi = f32c(lo, 3), f32c(hi, 3)
        check(self._lib.fp_flock_set_grid_domain(self._h, ptr(lo), ptr(hi)))

    def grid_info(self):
        dims = np.zeros(3, np.uint32)
        cell, bits = C.c_float(0), C.c_uint32(0)
        check(self._lib.fp_flock_grid_info(self._h, ptr(dims), C.byref(cell), C.byref(bits)))
        return dims, cell.value, bits.value

    def shard_info(self):
        """-> (rank, world, peer_mapped)"""
        a, b, c = C.c_int(0), C.c_int(0), C.c_int(0)
        check(self._lib.fp_flock_shard_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, bool(c.value)

    def set_rebin(self, skin: float = -1.0, plan_scale: float = 1.0) -> None:
        check(self._lib.fp_flock_set_rebin(self._h, skin, plan_scale))

    def rebin_info(self):
        """-> (skin, grid_steps, rebins, replayed): lazy re-binning of the grid path."""
        skin = C.c_float(0)
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        check(self._lib.fp_flock_rebin_info(self._h, C.byref(skin), C.byref(a), C.byref(b), C.byref(c)))
        return skin.value, a.value, b.value, c.value

    def timing_begin(self) -> None:
        check(self._lib.fp_flock_timing_begin(self._h))

    def timing_end(self):
        """-> (steps, span_ms, sort_ms, influence_ms) measured with CUDA events on the
        library's stream since ``timing_begin``."""
        n, t, s, w = C.c_uint32(0), C.c_float(0), C.c_float(0), C.c_float(0)
        check(self._lib.fp_flock_timing_end(self._h, C.byref(n), C.byref(t), C.byref(s), C.byref(w)))
        return n.value, t.value, s.value, w.value

    def state_euler(self, h: float) -> None:
        """``State::<boid>::euler_step(h)`` (state.rs:75-83) with frozen acceleration."""
        check(self._lib.fp_flock_state_euler(self._h, float(f32(h))))

    def state_rk4(self, h: float) -> None:
        check(self._lib.fp_flock_state_rk4(self._h, float(f32(h))))

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.fp_flock_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# The three lead-boid closures of the demo scene, demos/flocking.rs:105-107, :139-148
def demo_path_0(t):
    return (f32(25.0) * cosf(f32(t) / f32(12.0)), f32(0.5), f32(0.0))


def demo_path_1(t):
    t = f32(t)
    return (f32(15.0) * cosf(t / f32(12.0)), f32(6.0) + f32(5.0) * cosf(t / f32(12.0)),
            f32(15.0) * sinf(t / f32(12.0)))


def demo_path_2(t):
    t = f32(t)
    return (f32(25.0) * cosf(t / f32(10.0)), f32(1.0), f32(10.0) * sinf(t / f32(9.0)))


DEMO_PATHS = (demo_path_0, demo_path_1, demo_path_2)


def demo_simulation(which: int = 1, *, seed: int = synth.SEED, device: int = 0,
                    method: int = _lib.METHOD_AUTO) -> Simulation:
    """The two simulations of ``demos/flocking.rs:92-156`` (config C1)."""
    scene = synth.DEMO_SIM1 if which == 1 else synth.DEMO_SIM2
    obstacles = [Obstacle(o[:3], float(o[3])) for o in synth.DEMO_OBSTACLES]
    leads = [LeadBoid(DEMO_PATHS[k]) for k in scene["lead_paths"]]
    return Simulation(scene["spawn"], scene["num_boids"], None, leads, obstacles, None, seed=seed,
                      device=device, method=method)
