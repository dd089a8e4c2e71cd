"""Multi-GPU parity check, launched by tests/test_gpu_multi.py (or by hand) as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tests/mgpu_check.py
Every rank runs the same sharded flock; rank 0 compares against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.distributed as dist

from feriphys_b200 import _lib, synth
from feriphys_b200.flocking import BoundingBox, Obstacle, PointAttractor
from feriphys_b200.sharded import ShardedSimulation
from gpu_util import TABLES, FixedLead, bits, py_config, rel_err
from oracle_lib import Scene, oracle

f32 = np.float32


def make(st, method, c, tables, device, numerics=_lib.NUMERICS_EXACT):
    t = tables or {}
    sim = ShardedSimulation.from_global_state(
        st, dist,
        bounding_box=BoundingBox(t["bbox"][0:2], t["bbox"][2:4], t["bbox"][4:6]) if "bbox" in t else None,
        lead_boids=[FixedLead(r) for r in t["leads"]] if "leads" in t else None,
        obstacles=[Obstacle(o[:3], float(o[3])) for o in t["obstacles"]] if "obstacles" in t else None,
        attractors=[PointAttractor(a[:3], float(a[3])) for a in t["attractors"]] if "attractors" in t else None,
        method=method, device=device, numerics=numerics)
    sim.set_config(py_config(c))
    scene = Scene(leads=t.get("leads"), attractors=t.get("attractors"), obstacles=t.get("obstacles"),
                  bbox=t.get("bbox"))
    return sim, scene


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    orc = oracle()
    nt = max(1, (os.cpu_count() or 1) // world)
    c = orc.default_config()
    if os.environ.get("MGPU_ONLY") in ("lazy", "fast"):
        (lazy_slabs if os.environ["MGPU_ONLY"] == "lazy" else fast_numerics)(orc, c, rank, world, local, nt)
        dist.barrier()
        dist.destroy_process_group()
        if rank == 0:
            print("MGPU_OK", flush=True)
        return

    # ---- all-pairs, all-gather sharded: bit-identical to the oracle -------------------------
    st = synth.uniform_flock(3001, 60.0, seed=81)          # 3001: ragged last rank
    sim, sc = make(st, _lib.METHOD_ALLPAIRS, c, TABLES, local)
    gc, gh = sim.read_neighbors()
    ga, gcomp = sim.read_accel(components=True)
    sim.step_many(20)
    got = sim.read_state()
    if rank == 0:
        rc, rh, _ = orc.neighbors_rows(c, st, threads=nt)
        assert np.array_equal(gc, rc) and np.array_equal(gh, rh), "all-pairs neighbour sets"
        ra, rcomp, _ = orc.accel_rows(c, sc, st, threads=nt)
        assert np.array_equal(bits(ga), bits(ra)) and np.array_equal(bits(gcomp), bits(rcomp))
        cur = st
        for _ in range(20):
            cur, _ = orc.step(c, sc, cur, threads=nt)
        assert np.array_equal(bits(got), bits(cur)), "all-pairs sharded trajectory not bit-exact"
        print(f"[mgpu x{world}] all-pairs all-gather: bit-exact over 20 steps", flush=True)
    assert sim.status() & ~3 == 0          # steering-panic bits are legitimate here (boids inside obstacles)

    # ---- grid, x-slab sharded: neighbour sets bit-exact, accelerations 1e-5 ------------------
    n = 60000
    st = synth.uniform_flock(n, 340.0, seed=82)
    st[:, 3:] *= f32(40.0)                                 # fast boids: plenty of slab crossings
    c2 = orc.default_config(dt=0.004)
    tables = dict(TABLES, bbox=np.array([-60, 400, -60, 400, -60, 400], f32))   # walls outside the flock
    sim, sc = make(st, _lib.METHOD_GRID, c2, tables, local)
    gc, gh = sim.read_neighbors()
    ga = sim.read_accel()
    idx0, _ = sim.read_local()
    sim.step_many(30)
    got = sim.read_state()
    idx1, loc1 = sim.read_local()
    cen = sim.pair_census()
    flags = sim.status()
    counts = torch.tensor([len(idx0), len(idx1), len(np.setdiff1d(idx1, idx0))], device="cuda")
    dist.all_reduce(counts)
    if rank == 0:
        rc, rh, _ = orc.neighbors_rows(c2, st, threads=nt, grid=True)
        assert np.array_equal(gc, rc) and np.array_equal(gh, rh), "slab neighbour sets"
        ra, _, _ = orc.accel_rows(c2, sc, st, threads=nt, grid=True)
        assert rel_err(ga, ra) <= 1e-5, rel_err(ga, ra)
        cur = st
        for _ in range(30):
            cur, _ = orc.step(c2, sc, cur, threads=nt, grid=True)
        scale = np.maximum(1.0, np.linalg.norm(cur[:, :3], axis=1))
        err = (np.linalg.norm(got[:, :3] - cur[:, :3], axis=1) / scale).max()
        assert err <= 1e-4, err
        assert int(counts[0]) == n and int(counts[1]) == n, counts   # every boid owned exactly once
        assert int(counts[2]) > 0, "no boid migrated between slabs: the test is not exercising migration"
        rc2, _, _ = orc.neighbors_rows(c2, got, threads=nt, grid=True)
        assert int(cen[2]) == int(rc2.sum())
        print(f"[mgpu x{world}] grid slabs: neighbour sets exact, accel {rel_err(ga, ra):.1e}, "
              f"30-step err {err:.1e}, {int(counts[2])} migrations", flush=True)
    assert flags & ~3 == 0, flags           # no capacity / halo / slab-jump bits
    # local rows are really inside this rank's slab
    assert np.array_equal(bits(loc1), bits(got[idx1.astype(np.int64)]))

    # ---- switching partitions mid-run ----------------------------------------------------------
    sim.set_method(_lib.METHOD_ALLPAIRS)
    sim.step()
    sim.set_method(_lib.METHOD_GRID)
    sim.step()
    got2 = sim.read_state()
    if rank == 0:
        cur2 = got
        for _ in range(2):
            cur2, _ = orc.step(c2, sc, cur2, threads=nt, grid=True)
        assert np.abs(got2 - cur2).max() <= 1e-5 * max(1.0, float(np.abs(cur2).max()))
        print(f"[mgpu x{world}] partition switches ok", flush=True)
    lazy_slabs(orc, c, rank, world, local, nt)
    fast_numerics(orc, c, rank, world, local, nt)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_OK", flush=True)


def fast_numerics(orc, c, rank, world, local, nt):
    """FAST numerics on both partitions (neighbour sets exact, accelerations 1e-5, trajectories
    1e-4), and the host round trip fp_flock_read_local -> fp_flock_write_local."""
    st = synth.uniform_flock(9001, 110.0, seed=85)
    sim, sc = make(st, _lib.METHOD_ALLPAIRS, c, TABLES, local, _lib.NUMERICS_FAST)
    ga = sim.read_accel()
    sim.step_many(20)
    got = sim.read_state()
    if rank == 0:
        ra, _, _ = orc.accel_rows(c, sc, st, threads=nt, grid=True)
        assert rel_err(ga, ra) <= 1e-5, rel_err(ga, ra)
        cur = st
        for _ in range(20):
            cur, _ = orc.step(c, sc, cur, threads=nt, grid=True)
        scale = np.maximum(1.0, np.linalg.norm(cur[:, :3], axis=1))
        err = (np.linalg.norm(got[:, :3] - cur[:, :3], axis=1) / scale).max()
        assert err <= 1e-4, err
        print(f"[mgpu x{world}] FAST all-pairs all-gather: accel {rel_err(ga, ra):.1e}, 20-step err {err:.1e}",
              flush=True)
    n = 80000
    st = synth.uniform_flock(n, 380.0, seed=86)
    # (walls outside the flock: TABLES' own box ends at 200, in the middle of this one -- its 1 / d^2
    #  singularity shoots boids across a whole slab in one step once the slabs are 3 cell layers thin
    #  (8 ranks), which the library reports as FP_STATUS_SLAB_JUMP, as it should)
    tables = dict(TABLES, bbox=np.array([-60, 440, -60, 440, -60, 440], np.float32))
    sim, sc = make(st, _lib.METHOD_GRID, c, tables, local, _lib.NUMERICS_FAST)
    sim.set_rebin(skin=0.12)
    ga = sim.read_accel()
    gc, gh = sim.read_neighbors()
    sim.step_many(25)
    # host round trip of the rows each rank holds: values unchanged, the library re-bins
    idx, loc = sim.read_local()
    sim.write_local(idx, loc)
    sim.step_many(25)
    got = sim.read_state()
    skin, steps, rebins, replayed = sim.rebin_info()
    assert steps == 50 and rebins >= 3, (steps, rebins)
    if rank == 0:
        ra, _, _ = orc.accel_rows(c, sc, st, threads=nt, grid=True)
        assert rel_err(ga, ra) <= 1e-5, rel_err(ga, ra)
        rc, rh, _ = orc.neighbors_rows(c, st, threads=nt, grid=True)
        assert np.array_equal(gc, rc) and np.array_equal(gh, rh)
        cur = st
        for _ in range(50):
            cur, _ = orc.step(c, sc, cur, threads=nt, grid=True)
        scale = np.maximum(1.0, np.linalg.norm(cur[:, :3], axis=1))
        err = (np.linalg.norm(got[:, :3] - cur[:, :3], axis=1) / scale).max()
        assert err <= 1e-4, err
        print(f"[mgpu x{world}] FAST grid slabs + write_local round trip: accel {rel_err(ga, ra):.1e}, "
              f"50-step err {err:.1e}, {rebins} binnings", flush=True)
    assert sim.status() & ~3 == 0


def lazy_slabs(orc, c, rank, world, local, nt):
    """Grid slabs at everyday speeds: one binning serves many steps; in between the walk kernel
    pushes its boundary boids into the neighbours' ghost blocks over peer memory (or NCCL when
    FP_SHARD_PEER=0) and the mailbox is the step barrier."""
    n = 80000
    st = synth.uniform_flock(n, 380.0, seed=83)
    sim, sc = make(st, _lib.METHOD_GRID, c, None, local)
    sim.set_rebin(skin=0.12)     # a few binnings within 60 steps (the automatic skin of a flock this
    sim.step_many(7)             # small on two GPUs would make one binning last hundreds of steps)
    sim.step_many(53)
    got = sim.read_state()
    skin, steps, rebins, replayed = sim.rebin_info()
    _, _, peer = sim.shard_info()
    want_peer = os.environ.get("FP_SHARD_PEER", "1") != "0"
    gc, gh = sim.read_neighbors()
    idx, loc = sim.read_local()
    own = torch.tensor([len(idx)], device="cuda")
    dist.all_reduce(own)
    assert steps == 60 and replayed == 0 and 2 <= rebins <= 20, (steps, rebins, replayed)
    assert peer == want_peer, "peer mapping of the halo buffers is not in the expected state"
    assert np.array_equal(bits(loc), bits(got[idx.astype(np.int64)]))
    if rank == 0:
        cur = st
        for _ in range(60):
            cur, _ = orc.step(c, sc, cur, threads=nt, grid=True)
        scale = np.maximum(1.0, np.linalg.norm(cur[:, :3], axis=1))
        err = (np.linalg.norm(got[:, :3] - cur[:, :3], axis=1) / scale).max()
        assert err <= 1e-4, err
        assert int(own[0]) == n
        rc, rh, _ = orc.neighbors_rows(c, got, threads=nt, grid=True)
        assert np.array_equal(gc, rc) and np.array_equal(gh, rh), "slab neighbour sets after lazy steps"
        print(f"[mgpu x{world}] lazy slabs ({'peer stores' if peer else 'nccl messages'}): skin {skin:.3f}, "
              f"{rebins} binnings / 60 steps, 60-step err {err:.1e}, neighbour sets exact", flush=True)
    assert sim.status() == 0

    # a long run crosses the collective re-fit every 256 steps (slabs re-cut, peer buffers re-mapped)
    st3 = synth.uniform_flock(30000, 300.0, seed=84)
    sim, sc = make(st3, _lib.METHOD_GRID, c, None, local)
    sim.step_many(280)
    got = sim.read_state()
    gc, gh = sim.read_neighbors()
    assert sim.rebin_info()[1] == 280 and sim.status() == 0
    if rank == 0:
        cur = st3
        for _ in range(280):
            cur, _ = orc.step(c, sc, cur, threads=nt, grid=True)
        scale = np.maximum(1.0, np.linalg.norm(cur[:, :3], axis=1))
        err = (np.linalg.norm(got[:, :3] - cur[:, :3], axis=1) / scale).max()
        assert err <= 1e-4, err
        rc, rh, _ = orc.neighbors_rows(c, got, threads=nt, grid=True)
        assert np.array_equal(gc, rc) and np.array_equal(gh, rh), "slab neighbour sets after a re-fit"
        print(f"[mgpu x{world}] 280 steps across a re-fit: err {err:.1e}, neighbour sets exact", flush=True)

    # the plan is 50x too optimistic while an attractor speeds the flock up: every rank must void
    # the same step on the device and replay it after a fresh (collective) binning
    tables = dict(attractors=np.array([[-60, 190, 190, 2.0e4]], f32))
    sim, sc = make(st, _lib.METHOD_GRID, c, tables, local)
    sim.set_rebin(skin=0.1, plan_scale=50.0)
    sim.step_many(80)
    got = sim.read_state()
    skin, steps, rebins, replayed = sim.rebin_info()
    gc, gh = sim.read_neighbors()
    assert steps == 80 and replayed > 0, (steps, replayed)
    if rank == 0:
        cur = st
        for _ in range(80):
            cur, _ = orc.step(c, sc, cur, threads=nt, grid=True)
        scale = np.maximum(1.0, np.linalg.norm(cur[:, :3], axis=1))
        err = (np.linalg.norm(got[:, :3] - cur[:, :3], axis=1) / scale).max()
        assert err <= 1e-4, err
        rc, rh, _ = orc.neighbors_rows(c, got, threads=nt, grid=True)
        assert np.array_equal(gc, rc) and np.array_equal(gh, rh), "slab neighbour sets after replays"
        print(f"[mgpu x{world}] outrun plan: {replayed} steps voided on the device and replayed, "
              f"{rebins} binnings, 80-step err {err:.1e}", flush=True)
    assert sim.status() == 0


if __name__ == "__main__":
    main()
