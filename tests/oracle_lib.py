"""ctypes binding of the CPU oracle (oracle/flock_oracle.c).

TEST INFRASTRUCTURE: imported only by tests/, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs -- never by the
product package ``feriphys_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libflock_oracle.so")

FLAG_STEER_NEGATIVE = 1
FLAG_STEER_NAN_OVF = 2


class OrcConfig(C.Structure):
    _fields_ = [
        ("dt", C.c_float),
        ("avoidance_factor", C.c_float),
        ("centering_factor", C.c_float),
        ("velocity_matching_factor", C.c_float),
        ("distance_weight_threshold", C.c_float),
        ("distance_weight_threshold_falloff", C.c_float),
        ("max_sight_angle", C.c_float),
        ("max_sight_angle_to_lead_boid", C.c_float),
        ("time_to_start_steering_secs", C.c_uint64),
        ("time_to_start_steering_nanos", C.c_uint32),
        ("steering_overrides", C.c_int32),
    ]


class OrcScene(C.Structure):
    _fields_ = [
        ("leads", C.c_void_p),
        ("n_leads", C.c_uint32),
        ("attractors", C.c_void_p),
        ("n_attractors", C.c_uint32),
        ("obstacles", C.c_void_p),
        ("n_obstacles", C.c_uint32),
        ("bbox", C.c_void_p),
    ]


PATH_FN = C.CFUNCTYPE(None, C.c_float, C.POINTER(C.c_float), C.c_void_p)
DERIV_FN = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_size_t, C.c_void_p)


def build_oracle(force: bool = False) -> str:
    src = [os.path.join(ORACLE_DIR, f) for f in ("flock_oracle.c", "next_oracle.c", "flock_oracle.h", "Makefile")]
    stale = (not os.path.exists(ORACLE_SO)) or any(
        os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)
    return ORACLE_SO


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Scene:
    """Host-side tables kept alive for the lifetime of the OrcScene struct."""

    def __init__(self, leads=None, attractors=None, obstacles=None, bbox=None):
        self.leads = None if leads is None else _f32(leads).reshape(-1, 7)
        self.attractors = None if attractors is None else _f32(attractors).reshape(-1, 4)
        self.obstacles = None if obstacles is None else _f32(obstacles).reshape(-1, 4)
        self.bbox = None if bbox is None else _f32(bbox).reshape(6)

    def struct(self) -> OrcScene:
        s = OrcScene()
        if self.leads is not None and len(self.leads):
            s.leads, s.n_leads = _ptr(self.leads), len(self.leads)
        if self.attractors is not None and len(self.attractors):
            s.attractors, s.n_attractors = _ptr(self.attractors), len(self.attractors)
        if self.obstacles is not None and len(self.obstacles):
            s.obstacles, s.n_obstacles = _ptr(self.obstacles), len(self.obstacles)
        if self.bbox is not None:
            s.bbox = _ptr(self.bbox)
        return s


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.orc_sight_angle.restype = C.c_float
        L.orc_distance.restype = C.c_float
        L.orc_duration_as_secs_f32.restype = C.c_float
        L.orc_duration_as_secs_f32.argtypes = [C.c_uint64, C.c_uint32]
        L.orc_duration_from_secs_f32.restype = C.c_uint32
        L.orc_duration_from_secs_f32.argtypes = [C.c_float, C.POINTER(C.c_uint64),
                                                 C.POINTER(C.c_uint32)]
        L.orc_mix64.restype = C.c_uint64
        L.orc_mix64.argtypes = [C.c_uint64]
        L.orc_acos_threshold.restype = C.c_float
        L.orc_acos_threshold.argtypes = [C.c_float]
        L.orc_acos_monotone_violations.restype = C.c_uint64
        L.orc_acos_monotone_violations.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_grid_build.restype = C.c_void_p
        L.orc_grid_build.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_grid_free.argtypes = [C.c_void_p]
        L.orc_lead_step.restype = C.c_uint32
        L.orc_lead_step.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_float, PATH_FN,
                                    C.c_void_p]
        L.orc_demo_path.argtypes = [C.c_int, C.c_float, C.c_void_p]
        L.orc_pair_accel.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_float] * 7 + [
            C.c_void_p]
        L.orc_attractor_accel.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        L.orc_bbox_accel.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_steering_accel.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.POINTER(C.c_uint32)]
        rows = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64,
                C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_accel_rows.argtypes = rows
        L.orc_grid_accel_rows.argtypes = [C.c_void_p] + rows
        L.orc_step.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_int]
        L.orc_grid_step.argtypes = L.orc_step.argtypes
        L.orc_neighbors_rows.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                         C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_uint32, C.c_int]
        L.orc_grid_neighbors_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                              C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                              C.c_int]
        L.orc_pair_census.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64,
                                      C.c_void_p, C.c_int]
        L.orc_state_euler.argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_void_p, C.c_void_p,
                                      C.c_void_p]
        L.orc_state_rk4.argtypes = L.orc_state_euler.argtypes
        L.orc_instances.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p]

    # ---- config ---------------------------------------------------------
    def default_config(self, **overrides) -> OrcConfig:
        cfg = OrcConfig()
        self.lib.orc_config_default(C.byref(cfg))
        for k, v in overrides.items():
            if not hasattr(cfg, k):
                raise AttributeError(k)
            setattr(cfg, k, v)
        return cfg

    # ---- pair function & pieces ----------------------------------------
    def pair_accel(self, self6, other_pos, other_vel, weight=1.0, cfg=None, theta=None):
        cfg = cfg or self.default_config()
        s, p, v = _f32(self6), _f32(other_pos), _f32(other_vel)
        out = np.zeros(3, np.float32)
        th = cfg.max_sight_angle if theta is None else theta
        self.lib.orc_pair_accel(_ptr(s), _ptr(p), _ptr(v), weight, cfg.avoidance_factor,
                                cfg.centering_factor, cfg.velocity_matching_factor,
                                cfg.distance_weight_threshold,
                                cfg.distance_weight_threshold_falloff, th, _ptr(out))
        return out

    def sight_angle(self, self6, other_pos) -> float:
        s, p = _f32(self6), _f32(other_pos)
        return float(self.lib.orc_sight_angle(_ptr(s), _ptr(p)))

    def distance(self, self6, other_pos) -> float:
        s, p = _f32(self6), _f32(other_pos)
        return float(self.lib.orc_distance(_ptr(s), _ptr(p)))

    def attractor_accel(self, attr4, pos3, mass=1.0):
        a, p = _f32(attr4), _f32(pos3)
        out = np.zeros(3, np.float32)
        self.lib.orc_attractor_accel(_ptr(a), _ptr(p), mass, _ptr(out))
        return out

    def bbox_accel(self, bbox6, pos3):
        b, p = _f32(bbox6), _f32(pos3)
        out = np.zeros(3, np.float32)
        self.lib.orc_bbox_accel(_ptr(b), _ptr(p), _ptr(out))
        return out

    def steering_accel(self, cfg, scene: Scene, boid6):
        b = _f32(boid6)
        out = np.zeros(3, np.float32)
        flags = C.c_uint32(0)
        sc = scene.struct()
        self.lib.orc_steering_accel(C.byref(cfg), C.byref(sc), _ptr(b), _ptr(out),
                                    C.byref(flags))
        return out, flags.value

    # ---- rows / step ----------------------------------------------------
    def accel_rows(self, cfg, scene: Scene | None, state6, i0=0, i1=None, threads=1,
                   grid=False):
        st = _f32(state6).reshape(-1, 6)
        n = len(st)
        i1 = n if i1 is None else i1
        r = i1 - i0
        total = np.zeros((r, 3), np.float32)
        comp = np.zeros((r, 5, 3), np.float32)
        flags = np.zeros(r, np.uint32)
        sc = (scene or Scene()).struct()
        if grid:
            g = self.lib.orc_grid_build(C.byref(cfg), n, _ptr(st))
            try:
                self.lib.orc_grid_accel_rows(g, C.byref(cfg), C.byref(sc), n, _ptr(st), i0, i1,
                                             _ptr(total), _ptr(comp), _ptr(flags), threads)
            finally:
                self.lib.orc_grid_free(g)
        else:
            self.lib.orc_accel_rows(C.byref(cfg), C.byref(sc), n, _ptr(st), i0, i1, _ptr(total),
                                    _ptr(comp), _ptr(flags), threads)
        return total, comp, flags

    def step(self, cfg, scene: Scene | None, state6, threads=1, grid=False):
        st = _f32(state6).reshape(-1, 6)
        out = np.empty_like(st)
        flags = np.zeros(len(st), np.uint32)
        sc = (scene or Scene()).struct()
        fn = self.lib.orc_grid_step if grid else self.lib.orc_step
        fn(C.byref(cfg), C.byref(sc), len(st), _ptr(st), _ptr(out), _ptr(flags), threads)
        return out, flags

    def neighbors_rows(self, cfg, state6, i0=0, i1=None, threads=1, list_cap=0, grid=False):
        st = _f32(state6).reshape(-1, 6)
        n = len(st)
        i1 = n if i1 is None else i1
        r = i1 - i0
        count = np.zeros(r, np.uint32)
        hsh = np.zeros(r, np.uint64)
        if grid:
            g = self.lib.orc_grid_build(C.byref(cfg), n, _ptr(st))
            try:
                self.lib.orc_grid_neighbors_rows(g, C.byref(cfg), n, _ptr(st), i0, i1,
                                                 _ptr(count), _ptr(hsh), threads)
            finally:
                self.lib.orc_grid_free(g)
            return count, hsh, None
        lst = np.full((r, list_cap), 0xFFFFFFFF, np.uint32) if list_cap else None
        self.lib.orc_neighbors_rows(C.byref(cfg), n, _ptr(st), i0, i1, _ptr(count), _ptr(hsh),
                                    _ptr(lst) if lst is not None else None, list_cap, threads)
        return count, hsh, lst

    def pair_census(self, cfg, state6, i0=0, i1=None, threads=1):
        st = _f32(state6).reshape(-1, 6)
        i1 = len(st) if i1 is None else i1
        out = np.zeros(3, np.uint64)
        self.lib.orc_pair_census(C.byref(cfg), len(st), _ptr(st), i0, i1, _ptr(out), threads)
        return out

    def mix64(self, j: int) -> int:
        return int(self.lib.orc_mix64(j))

    # ---- Duration / leads ----------------------------------------------
    def duration_from_secs_f32(self, x: float):
        s, n = C.c_uint64(0), C.c_uint32(0)
        f = self.lib.orc_duration_from_secs_f32(x, C.byref(s), C.byref(n))
        return f, s.value, n.value

    def duration_as_secs_f32(self, secs: int, nanos: int) -> float:
        return float(self.lib.orc_duration_as_secs_f32(secs, nanos))

    def demo_path(self, kind: int, t: float):
        out = np.zeros(3, np.float32)
        self.lib.orc_demo_path(kind, t, _ptr(out))
        return out

    def make_leads(self, kinds):
        """LeadBoid::new for the demo closures (boid.rs:35-44): pos = path(0), vel = 0,
        weight = 10."""
        leads = np.zeros((len(kinds), 7), np.float32)
        for k, kind in enumerate(kinds):
            leads[k, :3] = self.demo_path(kind, 0.0)
            leads[k, 6] = 10.0
        return leads, np.zeros(len(kinds), np.float32)

    def step_leads(self, leads, times, kinds, dt: float):
        """lead_boid.step(Duration::from_secs_f32(dt)) for each lead, in place."""
        for k, kind in enumerate(kinds):
            cb = PATH_FN(lambda t, out, ctx, kind=kind: self.lib.orc_demo_path(kind, t, out))
            t = C.c_float(float(times[k]))
            row = np.ascontiguousarray(leads[k])
            f = self.lib.orc_lead_step(_ptr(row), C.byref(t), dt, cb, None)
            assert f == 0
            leads[k] = row
            times[k] = t.value

    def lead_step_fn(self, lead7, t: float, dt: float, path):
        """Generic LeadBoid::step with a Python path callable t -> (x, y, z)."""
        def _cb(tt, out, ctx):
            p = path(tt)
            out[0], out[1], out[2] = float(p[0]), float(p[1]), float(p[2])
        cb = PATH_FN(_cb)
        row = _f32(lead7).copy()
        tc = C.c_float(t)
        f = self.lib.orc_lead_step(_ptr(row), C.byref(tc), dt, cb, None)
        return f, row, tc.value

    # ---- State ----------------------------------------------------------
    def _deriv(self, deriv):
        if isinstance(deriv, str):
            fn = getattr(self.lib, "orc_deriv_" + deriv if hasattr(self.lib, "orc_deriv_" + deriv)
                         else "orc_deriv_test_" + deriv)
            return C.cast(fn, C.c_void_p), None
        def _cb(s, ds, n, ctx):
            sv = np.ctypeslib.as_array(s, shape=(n,))
            d = np.asarray(deriv(sv.copy()), dtype=np.float32)
            for i in range(n):
                ds[i] = d[i]
        cb = DERIV_FN(_cb)
        return C.cast(cb, C.c_void_p), cb

    def state_euler(self, s, h, deriv):
        s = _f32(s)
        out = np.empty_like(s)
        fn, keep = self._deriv(deriv)
        self.lib.orc_state_euler(_ptr(s), s.size, h, fn, None, _ptr(out))
        return out

    def state_rk4(self, s, h, deriv):
        s = _f32(s)
        out = np.empty_like(s)
        fn, keep = self._deriv(deriv)
        self.lib.orc_state_rk4(_ptr(s), s.size, h, fn, None, _ptr(out))
        return out

    def sph_neighbors(self, pos3, k=8, s=0.1, mass=0.001):
        """sph/mod.rs:89-121 -> (index [n, k], count [n], density [n])"""
        p = _f32(pos3).reshape(-1, 3)
        n = len(p)
        idx = np.zeros((n, k), np.uint32)
        cnt = np.zeros(n, np.uint32)
        den = np.zeros(n, np.float32)
        self.lib.orc_sph_neighbors.argtypes = [C.c_uint64, C.c_void_p, C.c_uint32, C.c_float, C.c_float,
                                               C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib.orc_sph_neighbors(n, _ptr(p), k, s, mass, _ptr(idx), _ptr(cnt), _ptr(den))
        return idx, cnt, den

    # ---- misc -----------------------------------------------------------
    def acos_threshold(self, theta: float) -> float:
        return float(self.lib.orc_acos_threshold(theta))

    def acos_monotone_violations(self, lo_bits: int, hi_bits: int) -> int:
        return int(self.lib.orc_acos_monotone_violations(lo_bits, hi_bits))

    def instances(self, state6):
        st = _f32(state6).reshape(-1, 6)
        out = np.zeros((len(st), 8), np.float32)
        self.lib.orc_instances(len(st), _ptr(st), _ptr(out))
        return out


_ORACLE = None


def oracle() -> Oracle:
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE
