#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native flocking step.

A "step" is one ``Simulation::step`` (flocking.rs:97-131) over a synthetic
flock that is already resident in HBM.  Default workload: BASELINE.json
configs[3] (C4) -- 2^24 boids, U[0,2048)^3, radius-limited FOV-gated influence
on the uniform-grid path -- which fits one B200 and is the configuration the
metric "boid-steps/sec at 16M boids, 1/2/4/8 B200" is quoted on; with
``--gpus N`` the SAME flock is slab-sharded over N ranks (strong scaling).
Other SURVEY 8d configurations: ``--workload c1|c2|c3|c5``.

Prints ONE JSON line (rank 0).  ``--impl reference`` times the CPU oracle (a C
restatement of the reference's Rust loops: the reference cannot be built here,
no Rust toolchain) on the host cores instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

f32 = np.float32
PI = float(f32(3.14159274101257324))

# SURVEY.md 8d.  n, domain edge, method, config overrides, default timed steps
WORKLOADS = {
    "c1": dict(n=110, desc="demo scene sim 1 (demos/flocking.rs:92-121): 110 boids, 1 lead, ship obstacle",
               method="small", steps=1000),
    "c2": dict(n=100_000, extent=24.0, method="allpairs", steps=10,
               desc="100k boids U[0,24)^3, all-pairs, distance-gated only (max_sight_angle=pi)"),
    "c3": dict(n=1 << 20, extent=816.0, method="grid", steps=100,
               desc="2^20 boids U[0,816)^3, uniform grid, FOV pi/2, defaults"),
    "c4": dict(n=1 << 24, extent=2048.0, method="grid", steps=100,
               desc="2^24 boids U[0,2048)^3, uniform grid, FOV pi/2, defaults, x-slab sharded"),
    "c5": dict(n=1 << 22, extent=1296.0, method="grid", steps=100,
               desc="2^22 boids U[0,1296)^3, 8 Lissajous leads, 8 attractors/repellers, 16 obstacles, bbox"),
}

FP32_LANES = 148 * 128  # B200: 148 SMs x 128 FP32 lanes


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--method", default=None, choices=["grid", "allpairs", "small"])
    ap.add_argument("--n", type=int, default=None, help="override the flock size (debugging)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="cpu_baseline budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--numerics", default="fast", choices=["exact", "fast"],
                    help="arithmetic of the forces: fast (default here; the north star's own bars: neighbour sets "
                         "bit-exact, accelerations within 1e-5) or exact (the library's default: every operation "
                         "separately rounded); the other one is measured beside it (other_numerics)")
    ap.add_argument("--no-alt", action="store_true", help="skip the comparison run with the other numerics")
    ap.add_argument("--no-parity", action="store_true", help="skip the sampled parity check against the oracle")
    return ap.parse_args()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), float(p.get("sm_max_mhz", 1965.0)), "MEASURED_PEAKS.json (measured)"
    except Exception:
        return 6650.0, 1965.0, "B200_PROFILING.md fallback (6.65 TB/s)"


# ---- workload construction (host arrays shared by both arms) ----------------------
def lissajous(k):
    from feriphys_b200.flocking import cosf, sinf
    a, b, c = f32(300 + 40 * k), f32(7 + k), f32(5 + 2 * k)

    def path(t):
        t = f32(t)
        return (f32(648) + a * cosf(t / b), f32(648) + a * sinf(t / c), f32(648) + a * cosf(t / (b + c)))
    return path


def build_workload(name, n_override=None, first=0, count=None):
    from feriphys_b200 import synth
    w = dict(WORKLOADS[name])
    w["name"] = name
    if n_override:
        w["n"] = n_override
    n = w["n"]
    cfg = {}
    tables = {}
    if name == "c1":
        state = synth.spawn_flock(synth.DEMO_SIM1["spawn"], n)
        tables["obstacles"] = synth.DEMO_OBSTACLES
        w["lead_paths"] = "demo"
    else:
        cnt = n if count is None else count
        state = synth.uniform_flock(cnt, w["extent"], first=first)
    if name == "c2":
        # factors scaled by 110/N keep accelerations at demo magnitude (inside the GUI ranges)
        cfg = dict(max_sight_angle=PI, centering_factor=float(f32(0.1) * f32(110.0) / f32(n)),
                   velocity_matching_factor=float(f32(0.5) * f32(110.0) / f32(n)))
    if name == "c5":
        att, obs, bbox = synth.c5_tables(w["extent"])
        tables.update(attractors=att, obstacles=obs, bbox=bbox)
        w["lead_paths"] = "lissajous"
    w.update(state=state, cfg=cfg, tables=tables)
    return w


def make_leads(w):
    from feriphys_b200.flocking import DEMO_PATHS, LeadBoid
    kind = w.get("lead_paths")
    if kind == "demo":
        return [LeadBoid(DEMO_PATHS[0])]
    if kind == "lissajous":
        return [LeadBoid(lissajous(k)) for k in range(8)]
    return None


def py_config(cfg_over):
    from feriphys_b200.flocking import Config
    c = Config()
    for k, v in cfg_over.items():
        setattr(c, k, v)
    return c


# ---- clocks during the timed region -------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 6:
                continue
            try:
                sm.append(float(c[0]))
                mx.append(float(c[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), c[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---- CPU oracle legs ------------------------------------------------------------------
class OracleRunner:
    """Times the CPU oracle on a bounded sample of the workload's rows."""

    def __init__(self, w, threads):
        from oracle_lib import Scene, oracle
        import ctypes as C
        self.C = C
        self.orc = oracle()
        self.w = w
        self.threads = threads
        self.cfg = self.orc.default_config(**w["cfg"])
        leads = make_leads(w)
        t = w["tables"]
        self.scene = Scene(leads=np.stack([l.row() for l in leads]) if leads else None,
                           attractors=t.get("attractors"), obstacles=t.get("obstacles"),
                           bbox=t.get("bbox"))
        self.sc = self.scene.struct()
        self.state = np.ascontiguousarray(w["state"], np.float32)
        self.n = len(self.state)
        self.grid = None
        self.t_build = 0.0
        if w["method_resolved"] == "grid":
            t0 = time.perf_counter()
            self.grid = self.orc.lib.orc_grid_build(C.byref(self.cfg), self.n,
                                                    self.state.ctypes.data_as(C.c_void_p))
            self.t_build = time.perf_counter() - t0

    def rows(self, m):
        """accelerations of rows [0, m): the per-boid work of one step. -> seconds"""
        C = self.C
        out = np.empty((m, 3), np.float32)
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        t0 = time.perf_counter()
        if self.grid is not None:
            self.orc.lib.orc_grid_accel_rows(self.grid, C.byref(self.cfg), C.byref(self.sc), self.n,
                                             P(self.state), 0, m, P(out), None, None, self.threads)
        else:
            self.orc.lib.orc_accel_rows(C.byref(self.cfg), C.byref(self.sc), self.n, P(self.state), 0,
                                        m, P(out), None, None, self.threads)
        return time.perf_counter() - t0

    def pick_rows(self, seconds):
        m0 = min(self.n, 4096 if self.grid is not None else 8 * self.threads)
        t = max(self.rows(m0), 1e-4)
        m = int(min(self.n, max(m0, m0 * seconds / t)))
        return max(1, m)

    def close(self):
        if self.grid is not None:
            self.orc.lib.orc_grid_free(self.grid)
            self.grid = None


def literal_rows_rate(r, seconds):
    """The reference's own algorithm on one thread: literal O(N) rows of the oracle (no grid).
    -> {"value": boid-steps/s, "rows": m, "sample": ...}"""
    C = r.C
    P = lambda a: a.ctypes.data_as(C.c_void_p)

    def go(m):
        out = np.empty((m, 3), np.float32)
        t0 = time.perf_counter()
        r.orc.lib.orc_accel_rows(C.byref(r.cfg), C.byref(r.sc), r.n, P(r.state), 0, m, P(out), None, None, 1)
        return time.perf_counter() - t0

    m = 1
    t = max(go(m), 1e-6)
    if t < seconds / 4:
        m = int(max(1, min(r.n, seconds / t)))
        t = max(go(m), 1e-6)
    return {"value": m / t, "unit": "boid-steps/s", "cores": 1, "rows": m,
            "sample": f"accelerations of the first {m} of {r.n} boids, each against all {r.n} (the reference's "
                      f"O(N^2) loop, single thread, as the reference runs it)"}


def sample_text(r, m):
    how = "grid-accelerated oracle (bit-identical to the literal loops)" if r.grid is not None \
        else "literal O(N) rows"
    return (f"accelerations of the first {m} of {r.n} boids per step ({how}); "
            f"boid-steps/s = rows / seconds" +
            (f"; one-off grid build {r.t_build:.2f}s excluded" if r.grid is not None else ""))


def pairs_per_boid_step(census, n):
    """in-range ordered pairs per step (grid) or all ordered pairs (all-pairs)."""
    return float(census[1] + census[2]), float(census[3])


# ---- reference arm ---------------------------------------------------------------------
def run_reference(args, w, rank):
    if rank != 0:
        return None
    threads = os.cpu_count() or 1
    r = OracleRunner(w, threads)
    steps = args.steps
    per_step = min(3.0, 120.0 / max(1, steps + args.warmup))
    m = r.pick_rows(per_step)
    for _ in range(args.warmup):
        r.rows(m)
    t = 0.0
    for _ in range(steps):
        t += r.rows(m)
    sample = sample_text(r, m)
    r.close()
    value = m * steps / t
    line = {
        "impl": "reference", "metric": "boid-steps/sec", "value": value, "unit": "boid-steps/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (splitmix64-keyed uniform flock, seed 0xFE21F)",
        "config": config_block(w, args, 1),
        "cpu_baseline": {"value": value, "unit": "boid-steps/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "boid-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = CPU oracle (C restatement of the Rust loops, OpenMP over boids); "
                "the Rust reference cannot be built in this image (no cargo/rustc)",
    }
    return line


def config_block(w, args, world):
    return {
        "workload": f"{w['name']}: {w['desc']}",
        "boids": int(w["n"]), "method": w["method_resolved"],
        "sharding": ("none" if world == 1 else
                     (f"x-slab x{world}, halo exchange + migration" if w["method_resolved"] == "grid"
                      else f"boid-index x{world}, all-gather of pos/vel")),
        "l2": "inputs larger than L2 (state is %.0f MB per buffer)" % (w["n"] * 32 / 1e6)
        if w["n"] * 32 > 126e6 else "state fits L2; no flush (FP32/latency-bound workload)",
    }


# ---- our arm -----------------------------------------------------------------------------
def parity_check(sim, w, rank, rows=2048):
    """Sampled parity of the state the bench ends on, outside the timed region: the library's
    neighbour sets and accelerations (collective taps on a sharded flock) against the oracle run
    on the same state on rank 0.  -> dict for the JSON line (None on the other ranks)."""
    state = sim.read_state()                      # global state, caller order (every rank)
    n = len(state)
    grid = w["method_resolved"] == "grid" or n > 200_000
    half = max(1, min(rows, n) // 2)
    wins = [(0, half), (n - half, n)] if n > 2 * half else [(0, n)]
    gc, gh = sim.read_neighbors()
    ga = sim.read_accel()
    if rank != 0:
        return None
    import ctypes as C
    from oracle_lib import Scene, oracle
    orc = oracle()
    cfg = orc.default_config(**w["cfg"])
    t = w["tables"]
    leads = getattr(sim, "_lead_rows")()
    sc = Scene(leads=leads if len(leads) else None, attractors=t.get("attractors"),
               obstacles=t.get("obstacles"), bbox=t.get("bbox"))
    threads = os.cpu_count() or 1
    mism, worst, nrows, order_noise = 0, 0.0, 0, 0.0
    g = orc.lib.orc_grid_build(C.byref(cfg), n, state.ctypes.data_as(C.c_void_p)) if grid else None
    scs = sc.struct()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    try:
        for lo, hi in wins:
            m = hi - lo
            rc, rh = np.zeros(m, np.uint32), np.zeros(m, np.uint64)
            ra = np.zeros((m, 3), np.float32)
            if grid:
                orc.lib.orc_grid_neighbors_rows(g, C.byref(cfg), n, P(state), lo, hi, P(rc), P(rh), threads)
                orc.lib.orc_grid_accel_rows(g, C.byref(cfg), C.byref(scs), n, P(state), lo, hi, P(ra), None, None,
                                            threads)
            else:
                orc.lib.orc_neighbors_rows(C.byref(cfg), n, P(state), lo, hi, P(rc), P(rh), None, 0, threads)
                orc.lib.orc_accel_rows(C.byref(cfg), C.byref(scs), n, P(state), lo, hi, P(ra), None, None, threads)
            mism += int(np.count_nonzero((gc[lo:hi] != rc) | (gh[lo:hi] != rh)))
            num = np.linalg.norm(ga[lo:hi].astype(np.float64) - ra.astype(np.float64), axis=1)
            den = np.maximum(np.linalg.norm(ra.astype(np.float64), axis=1), 1e-3)
            worst = max(worst, float((num / den).max()))
            nrows += m
            if not grid and n >= 20_000:
                # A dense all-pairs flock sums tens of thousands of terms per boid: the reference's own
                # f32 sum moves when its loop order changes.  Measure by how much (the same rows with
                # the other boids listed in reverse) -- a kernel that sums in another order again
                # (FAST numerics: j split across lanes) cannot be held to less than that.
                perm = np.concatenate([np.arange(lo, hi), np.arange(lo - 1, -1, -1), np.arange(n - 1, hi - 1, -1)])
                sp = np.ascontiguousarray(state[perm])
                rb = np.zeros((m, 3), np.float32)
                orc.lib.orc_accel_rows(C.byref(cfg), C.byref(scs), n, P(sp), 0, m, P(rb), None, None, threads)
                nz = np.linalg.norm(rb.astype(np.float64) - ra.astype(np.float64), axis=1)
                order_noise = max(order_noise, float((nz / den).max()))
    finally:
        if g is not None:
            orc.lib.orc_grid_free(g)
    out = {"rows": nrows, "neighbor_mismatches": mism, "max_rel_accel": worst, "accel_bar": 1e-5,
           "oracle": "grid-accelerated (bit-identical to the literal loops)" if grid else "literal O(N) rows",
           "state": "the state the run ended on (after warm-up, timed steps and the e2e legs)"}
    if order_noise > 0.0:
        out["reference_order_sensitivity"] = order_noise
        out["accel_bar"] = max(1e-5, 2.0 * order_noise)
        out["accel_bar_note"] = ("max(1e-5, 2 x how far the reference's own f32 sums move when its loop over the "
                                 "other boids runs in reverse) -- tests/test_gpu_scale.py uses the same bar")
    return out


def timed_windows(sim, K, W, barrier, reduce_max, grid, reset, segment_steps=300, budget_s=25.0):
    """K-step windows, each bracketed by a barrier and timed with CUDA events on the library's
    stream (max over ranks), repeated until >= 1 s of device time has been timed and -- on the grid
    path -- >= 3 binnings fell inside.  The reference's flock accelerates without bound (no drag,
    flocking.rs:116-117), so a long run drifts away from the workload SURVEY 8d defines (the first
    few hundred steps of the synthetic state): every `segment_steps` timed steps the flock is put
    back to its initial state and warmed up again (W steps, untimed).  -> (windows, totals)"""
    wins, tot_s, tot_steps, bins, sort_ms, infl_ms, seen, replay = [], 0.0, 0, 0, 0.0, 0.0, 0, 0
    t_start = time.perf_counter()
    since_reset, resets = 0, 0
    while True:
        if since_reset and since_reset + K > max(segment_steps, K):
            reset()
            if W:
                sim.step_many(W)
            sim.sync()
            since_reset = 0
            resets += 1
        rb0 = sim.rebin_info() if grid else None
        barrier()
        sim.timing_begin()
        t0 = time.perf_counter()
        sim.step_many(K)
        nst, span_ms, so, inf = sim.timing_end()          # synchronises this rank's stream
        wall = time.perf_counter() - t0
        rb1 = sim.rebin_info() if grid else None
        dev_s, wall_s = reduce_max(span_ms / 1e3, wall)
        b = (rb1[2] - rb0[2]) if grid else 0
        wins.append({"ms_per_step": 1e3 * dev_s / K, "binnings": int(b), "wall_ms": 1e3 * wall_s,
                     "device_span_ms_this_rank": span_ms})
        tot_s += dev_s
        tot_steps += K
        since_reset += K
        bins += b
        sort_ms += so
        infl_ms += inf
        seen += max(1, nst)
        replay += (rb1[3] - rb0[3]) if grid else 0
        enough = tot_s >= 1.0 and (not grid or bins >= 3)
        if enough or len(wins) >= 2000 or time.perf_counter() - t_start > budget_s:
            break
    return wins, dict(dev_s=tot_s, steps=tot_steps, binnings=int(bins), sort_ms=sort_ms, infl_ms=infl_ms,
                      steps_seen=seen, replayed=int(replay), resets=resets)


def run_ours(args, w, rank, world, local_rank):
    from ctypes import byref, c_uint64 as C_uint64

    import torch
    from feriphys_b200 import _lib
    from feriphys_b200.flocking import Simulation, Obstacle, PointAttractor, BoundingBox
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(*vals):
        if dist is None:
            return vals
        tt = torch.tensor(list(vals), device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tuple(float(x) for x in tt)

    method = {"grid": _lib.METHOD_GRID, "allpairs": _lib.METHOD_ALLPAIRS,
              "small": _lib.METHOD_SMALL}[w["method_resolved"]]
    t = w["tables"]
    numerics = {None: None, "exact": _lib.NUMERICS_EXACT, "fast": _lib.NUMERICS_FAST}[args.numerics]
    kw = dict(
        bounding_box=(BoundingBox(t["bbox"][0:2], t["bbox"][2:4], t["bbox"][4:6]) if "bbox" in t else None),
        lead_boids=make_leads(w),
        obstacles=[Obstacle(o[:3], float(o[3])) for o in t["obstacles"]] if "obstacles" in t else None,
        attractors=[PointAttractor(a[:3], float(a[3])) for a in t["attractors"]] if "attractors" in t else None,
        method=method, device=local_rank, numerics=numerics)
    n = w["n"]
    if world == 1:
        sim = Simulation.from_state(w["state"], **kw)
    else:
        from feriphys_b200.sharded import ShardedSimulation
        sim = ShardedSimulation.from_global_slice(w["state"], n, w["first"], dist, **kw)
    sim.set_config(py_config(w["cfg"]))
    numerics_in_use = "fast" if sim.numerics()[1] == _lib.NUMERICS_FAST else "exact"
    grid = w["method_resolved"] == "grid"

    K, W = args.steps, args.warmup
    lib = _lib.load()
    sim.step_many(W) if W else None
    sim.sync()
    census = sim.pair_census()          # outside the timed region (extra launches)
    sampler = ClockSampler(local_rank)
    sampler.start()                     # (before the barrier: the Popen must not sit inside a window)
    from feriphys_b200 import synth

    def reset():
        """the flock back to its initial state (each rank: the boids it holds now), leads restarted"""
        if world == 1:
            sim.write_state(w["state"])
        else:
            idx, _ = sim.read_local()
            sim.write_local(idx, synth.uniform_flock(0, w["extent"], index=idx))
        if sim.lead_boids:
            sim.lead_boids = make_leads(w)
            sim._push_leads()

    launches0 = lib.fp_launch_count()
    wins, tot = timed_windows(sim, K, W, barrier, reduce_max, grid, reset)
    launches = lib.fp_launch_count() - launches0
    dev_s = tot["dev_s"]
    if dist is not None:
        ll = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(ll)
        launches = int(ll[0])
        # sanity of the sharded run: every boid owned by exactly one rank, no capacity / halo /
        # slab-jump / barrier flags raised on any rank
        own = C_uint64()
        _lib.check(lib.fp_flock_local_len(sim._h, byref(own)))
        bad = int(sim.status()) & ~3        # (bits 1 | 2: steering terms the reference would panic on)
        chk = torch.tensor([int(own.value), 1 if bad else 0], device="cuda", dtype=torch.int64)
        dist.all_reduce(chk)
        if int(chk[0]) != n or int(chk[1]):
            raise SystemExit(f"sharded run inconsistent: {int(chk[0])} boids owned of {n}, "
                             f"{int(chk[1])} ranks raised capacity/halo/barrier flags (this rank: {bad})")
    clocks = sampler.stop()
    steps_timed = tot["steps"]
    value = n * steps_timed / dev_s
    ms_per_step = 1e3 * dev_s / steps_timed

    # ---- e2e: the drop-in call sequence with HOST buffers, copies inside the timed region.  Every rank
    # uploads the rows it holds from pinned memory, steps once, reads them back -- the same at every N.
    e2e = None
    if not args.no_e2e:
        reset()
        ke = max(1, min(K, 10))
        if world == 1:
            host_in = torch.from_numpy(np.ascontiguousarray(w["state"])).pin_memory()
            host_out = torch.empty_like(host_in).pin_memory()
            a_in, a_out = host_in.numpy(), host_out.numpy()
            for _ in range(2):                                                     # warm (both upload slots)
                sim.write_state(a_in); sim.step(); a_out[:] = sim.read_state()
            barrier()
            te = time.perf_counter()
            for _ in range(ke):
                _lib.check(lib.fp_flock_write_state(sim._h, _lib.ptr(a_in)))     # H2D
                sim.step()                                                        # Simulation::step
                _lib.check(lib.fp_flock_read_state(sim._h, _lib.ptr(a_out)))     # D2H (synchronises)
            barrier()
            te = time.perf_counter() - te
            e2e = {"value": n * ke / te, "unit": "boid-steps/s", "h2d_bytes_per_step": int(n * 24),
                   "d2h_bytes_per_step": int(n * 24), "steps": ke,
                   "path": "fp_flock_write_state -> fp_flock_step -> fp_flock_read_state, pinned host buffers"}
            # the reference's own per-frame call sequence (demos/flocking.rs:215-227): step(), then
            # get_boid_instances(); the state stays where Simulation keeps it
            inst = torch.empty((n, 8), dtype=torch.float32).pin_memory().numpy()
            _lib.check(lib.fp_flock_read_instances(sim._h, _lib.ptr(inst)))
            barrier()
            tf = time.perf_counter()
            for _ in range(ke):
                sim.step()
                _lib.check(lib.fp_flock_read_instances(sim._h, _lib.ptr(inst)))
            barrier()
            tf = time.perf_counter() - tf
            e2e["frame"] = {"value": n * ke / tf, "unit": "boid-steps/s", "d2h_bytes_per_step": int(n * 32),
                            "h2d_bytes_per_step": int(28 * len(sim.lead_boids or [])),
                            "path": "Simulation.step() -> get_boid_instances() each step (lead rows up, instances down)"}
        else:
            cap_l = int(min(n, 2 * (n // world) + (1 << 20)))       # slabs are never this unbalanced here
            idx_l = torch.empty(cap_l, dtype=torch.int64).pin_memory().numpy().view(np.uint64)
            st_l = torch.empty((cap_l, 6), dtype=torch.float32).pin_memory().numpy()
            own = C_uint64()

            def rows():
                _lib.check(lib.fp_flock_local_len(sim._h, byref(own)))
                assert own.value <= cap_l
                return int(own.value)

            sim.step_many(1)
            m = rows()
            _lib.check(lib.fp_flock_read_local(sim._h, _lib.ptr(idx_l), _lib.ptr(st_l)))    # warm
            h2d = d2h = 0
            barrier()
            te = time.perf_counter()
            for _ in range(ke):
                _lib.check(lib.fp_flock_write_local(sim._h, m, _lib.ptr(idx_l), _lib.ptr(st_l)))   # H2D
                sim.step_many(1)
                h2d += m * 32
                m = rows()                                                                    # (boids migrate)
                _lib.check(lib.fp_flock_read_local(sim._h, _lib.ptr(idx_l), _lib.ptr(st_l)))   # D2H
                d2h += m * 32
            barrier()
            te = time.perf_counter() - te
            tt = torch.tensor([te], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            bb = torch.tensor([h2d, d2h], device="cuda", dtype=torch.int64)
            dist.all_reduce(bb)
            e2e = {"value": n * ke / float(tt[0]), "unit": "boid-steps/s",
                   "h2d_bytes_per_step": int(bb[0]) // ke, "d2h_bytes_per_step": int(bb[1]) // ke, "steps": ke,
                   "path": "every rank: fp_flock_write_local (the rows it holds, pinned) -> fp_flock_step -> "
                           "fp_flock_read_local; bytes summed over ranks"}

    # ---- the other arithmetic, for comparison (same flock, same harness, shorter)
    other = None
    if not args.no_alt and w["method_resolved"] != "small":
        alt = _lib.NUMERICS_EXACT if numerics_in_use == "fast" else _lib.NUMERICS_FAST
        sim.set_numerics(alt)
        if (sim.numerics()[1] == _lib.NUMERICS_FAST) != (numerics_in_use == "fast"):
            reset()
            sim.step_many(max(3, W))
            sim.sync()
            awins, atot = timed_windows(sim, K, W, barrier, reduce_max, grid, reset, budget_s=8.0)
            other = {"numerics": "exact" if alt == _lib.NUMERICS_EXACT else "fast",
                     "ms_per_step": 1e3 * atot["dev_s"] / atot["steps"], "value": n * atot["steps"] / atot["dev_s"],
                     "influence_ms_per_step": atot["infl_ms"] / atot["steps_seen"], "steps": atot["steps"],
                     "binnings": atot["binnings"]}
        sim.set_numerics(_lib.NUMERICS_FAST if numerics_in_use == "fast" else _lib.NUMERICS_EXACT)
        reset()
        sim.step_many(max(3, W))

    parity = None if args.no_parity else parity_check(sim, w, rank)

    hbm, sm_max, peak_src = measured_peaks()
    steps_seen = max(1, tot["steps_seen"])
    in_range, candidates = pairs_per_boid_step(census, n)
    # (on a sharded flock fp_flock_pair_census already returns the all-rank totals)
    flops_step = 8.0 * float(census[0]) + 18.0 * float(census[1]) + 54.0 * float(census[2])
    infl_s = tot["infl_ms"] / 1e3 / steps_seen if w["method_resolved"] != "small" else dev_s / steps_timed
    n_local = n / world
    fma_peak = FP32_LANES * 2 * sm_max * 1e6 / 1e12
    fast = numerics_in_use == "fast"
    # The binder of every influence kernel here is FP32 instruction issue, not HBM and not the tensor
    # cores (the path is no contraction).  EXACT numerics forbid FMA contraction: their cap is the
    # non-FMA issue peak, half of the FMA peak FAST numerics are measured against.
    peak = fma_peak if fast else fma_peak / 2
    roofline = {"bound": "fp32_issue", "achieved": flops_step / world / infl_s / 1e12, "peak": peak, "unit": "TFLOP/s",
                "flops_per_step": flops_step,
                "peak_source": f"148 SM x 128 FP32 lanes x clocks.max.sm ({sm_max:.0f} MHz): "
                               + ("x 2 (FMA peak)" if fast else "x 1 (exact arithmetic cannot fuse: non-FMA issue peak)"),
                "model": "8 / 18 / 54 flops per examined pair by outcome (SURVEY 8d.1), pairs counted once on the "
                         "post-warm-up state, over the influence kernel's CUDA-event time",
                "traffic": None}
    roofline["frac"] = roofline["achieved"] / roofline["peak"]
    if grid:
        lists = os.environ.get("FP_NL", "1") != "0"
        roofline["kernel"] = (("nl_fast_kernel" if fast else "nl_walk_kernel") + " (walk on standing candidate lists"
                              " + extras + Euler)") if lists else "grid_walk3_kernel<TAP_STEP> (TMA-staged 27-cell walk)"
        traffic, traffic_src = None, None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` launch (profiles/)
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
                tr = json.load(fh).get(f"{w['name']}_{numerics_in_use}" if lists else w["name"])
            if tr and tr["boids"] == n and world == 1:
                traffic, traffic_src = tr["dram_bytes_per_launch"], tr["source"]
        except Exception:
            pass
        roofline["traffic"] = traffic
        roofline["traffic_source"] = traffic_src
        # the HBM figure the north star names, beside the real binder
        roofline["hbm"] = {"achieved": 64.0 * n_local / infl_s / 1e9, "peak": hbm, "unit": "GB/s",
                           "frac": 64.0 * n_local / infl_s / 1e9 / hbm, "algorithmic_bytes_per_boid": 64,
                           "algorithmic_bytes_per_launch": 64 * int(n_local), "peak_source": peak_src,
                           "note": "SURVEY 8d.2 floor of 64 B/boid (state read + write) over the walk's time; the "
                                   "walk also streams its candidate lists (~90 B/boid) and the SoA copy (12 B)"}
    else:
        roofline["kernel"] = ("allpairs_fast_kernel (j split across lanes, warp-shuffle reduction)" if fast else
                              "allpairs2_kernel / allpairs_kernel<TAP_STEP> (the faster of the two, measured by the "
                              "handle)") if w["method_resolved"] == "allpairs" else "small_kernel"
    line = {
        "metric": "boid-steps/sec", "value": value, "unit": "boid-steps/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (splitmix64-keyed uniform flock, seed 0xFE21F)",
        "config": dict(config_block(w, args, world), numerics=numerics_in_use + (
            " (neighbour sets bit-exact; forces fused, accelerations within 1e-5 of the reference)" if fast else
            " (every operation separately rounded in the reference's order)")),
        "pair_interactions_per_sec": in_range * steps_timed / dev_s,
        "candidate_pairs_per_sec": candidates * steps_timed / dev_s,
        "pairs": {"rejected_by_distance": float(census[0]), "fov_culled": float(census[1]),
                  "contributing": float(census[2]), "examined": float(census[3]),
                  "note": "ordered pairs per step, counted once on the post-warm-up state"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "timing": {"windows": len(wins), "steps_timed": steps_timed, "device_s_timed": dev_s,
                   "state_resets": tot["resets"],
                   "window_ms_per_step": {"median": float(np.median([x["ms_per_step"] for x in wins])),
                                          "min": min(x["ms_per_step"] for x in wins),
                                          "max": max(x["ms_per_step"] for x in wins)},
                   "windows_head": wins[:12],
                   "sort_phase_ms_per_step": tot["sort_ms"] / steps_seen,
                   "influence_ms_per_step": tot["infl_ms"] / steps_seen,
                   "how": f"windows of {K} steps, each bracketed by a barrier and timed with CUDA events on the "
                          "library's stream, max over ranks; repeated until >= 1 s of device time and (grid) >= 3 "
                          "binnings were inside; value and ms_per_step are totals over all windows, so every "
                          "binning is paid for; sort_phase is amortised over the binnings observed; the flock is put "
                          "back to its initial state (+ warm-up, untimed) every 300 timed steps so that the "
                          "workload stays the one SURVEY 8d defines (the reference's flock accelerates for ever)"},
        "roofline": roofline, "other_numerics": other, "parity": parity,
    }
    if grid:
        rbi = sim.rebin_info()
        line["rebinning"] = {"skin": rbi[0], "binnings_in_timed_steps": tot["binnings"],
                             "steps_per_binning": steps_timed / max(1, tot["binnings"]),
                             "steps_replayed": tot["replayed"],
                             "halo": (("peer stores fused into the walk kernel (cudaIpc over NVLink), mailbox "
                                       "step barrier" if sim.shard_info()[2] else "ncclSend/ncclRecv after each walk")
                                      if world > 1 else None),
                             "note": "lazy re-binning: one sort by cell (+ candidate-list build) serves every step "
                                     "until some boid could have moved skin/2 (device-checked)"}
        step_s = dev_s / steps_timed
        # 64 (walk) + 68 (gather) + 16 * 3 (sort passes) + 16 (keys) = 196 B per boid and
        # binning; with lazy re-binning only the walk's 64 B recur every step
        bytes_step = 64.0 + 132.0 * tot["binnings"] / steps_timed
        line["roofline_step"] = {"bound": "hbm", "achieved": bytes_step * n_local / step_s / 1e9, "peak": hbm,
                                 "unit": "GB/s", "frac": bytes_step * n_local / step_s / 1e9 / hbm,
                                 "algorithmic_bytes_per_boid_step": bytes_step,
                                 "note": "whole step, SURVEY 8d.2 byte model: 64 B per step + 132 B (keys, sort, "
                                         "gather) per binning x binnings per step in the timed region"}
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        r = OracleRunner(w, threads)
        m = r.pick_rows(args.cpu_seconds)
        tsec = r.rows(m)
        line["cpu_baseline"] = {"value": m / tsec, "unit": "boid-steps/s", "cores": threads, "kind": "port",
                                "sample": sample_text(r, m)}
        # beside it, what the reference itself does (flocking.rs:133-151): one thread, every other boid
        # visited for every boid -- a few literal rows, ~2 s
        line["cpu_baseline"]["literal_single_thread"] = literal_rows_rate(r, 2.0)
        r.close()
    else:
        line["cpu_baseline"] = None
    if dist is not None:
        dist.destroy_process_group()
    return line if rank == 0 else None


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.impl == "reference" and rank != 0:
        return
    n = args.n or WORLDS_N(args.workload)
    first, count = 0, None
    if world > 1 and args.impl == "ours":
        from feriphys_b200.sharded import shard_range
        first, count = shard_range(n, rank, world)
    w = build_workload(args.workload, args.n, first, count)
    w["first"] = first
    w["method_resolved"] = args.method or w["method"]
    if args.steps is None:
        args.steps = w["steps"]
    if args.impl == "reference":
        line = run_reference(args, w, rank)
    else:
        line = run_ours(args, w, rank, world, local_rank)
    if line is not None:
        print(json.dumps(line), flush=True)


def WORLDS_N(name):
    return WORKLOADS[name]["n"]


if __name__ == "__main__":
    main()
