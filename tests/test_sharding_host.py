"""CPU, world_size 2 over gloo: the host-side plumbing of the sharded path --
index partition, unique-id broadcast, SPMD construction arguments."""
import os
import subprocess
import sys

import numpy as np

from feriphys_b200.sharded import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["FP_ROOT"])
import torch.distributed as dist
from feriphys_b200 import synth
from feriphys_b200.sharded import broadcast_unique_id, shard_range
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
uid = broadcast_unique_id(dist)
# every rank ends with rank 0's id
import torch
t = torch.from_numpy(uid.astype(np.int64))
g = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(g, t)
assert all(torch.equal(g[0], x) for x in g) and int(uid.astype(bool).sum()) > 4
# ranks generate disjoint pieces of one global flock straight from the keyed generator
n = 1001
first, count = shard_range(n, rank, world)
mine = synth.uniform_flock(count, 50.0, seed=3, first=first)
full = synth.uniform_flock(n, 50.0, seed=3)
assert np.array_equal(mine, full[first:first + count])
tot = torch.tensor([count]); dist.all_reduce(tot); assert int(tot) == n
dist.barrier(); dist.destroy_process_group()
print("HOST_OK", rank)
'''


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 1000, 1001, 1 << 24):
        for world in (1, 2, 3, 4, 8):
            rows = [shard_range(n, r, world) for r in range(world)]
            per = (n + world - 1) // world
            assert sum(c for _, c in rows) == n
            pos = 0
            for r, (first, count) in enumerate(rows):
                assert first == min(r * per, n) and count <= per
                assert first == pos or count == 0
                pos += count


def test_gloo_world2_bootstrap(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, FP_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29519", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.count("HOST_OK") == 2, r.stdout[-2000:] + r.stderr[-2000:]


# ---- the slab protocol of fp_shard.cu, emulated with numpy over gloo ------------------------
SLAB_WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["FP_ROOT"])
import torch, torch.distributed as dist
from scipy.spatial import cKDTree
from feriphys_b200 import synth
f32 = np.float32
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()

REACH, SKIN, ZSPAN = 16.0, 0.5, 2
n = 24000
st = synth.uniform_flock(n, 200.0, seed=11)               # every rank knows the global flock
pos = st[:, :3].copy()
lo, hi = pos.min(axis=0), pos.max(axis=0)
cell = REACH * (1 + 1 / 512) + SKIN
gdim = [int(np.floor((float(hi[a]) - float(lo[a])) / cell)) + 1 for a in range(3)]
dimz = int(np.floor((float(hi[2]) - float(lo[2])) / (cell / ZSPAN))) + 1
inv, invz = f32(1) / f32(cell), f32(ZSPAN / cell)
def coord(x, o, iv, d):
    return np.clip(np.floor((x.astype(f32) - f32(o)).astype(f32) * f32(iv)).astype(np.int64), 0, d - 1)
xs0, xs1 = gdim[0] * rank // world, gdim[0] * (rank + 1) // world     # owned global x layers
ldim0 = xs1 - xs0 + 2                                                 # + one ghost layer each side
CL = gdim[1] * dimz                                                   # cells per x layer
def keys_of(p):                                                       # x slowest, z fastest
    gx = coord(p[:, 0], lo[0], inv, gdim[0])
    cx = np.clip(gx - (xs0 - 1), 0, ldim0 - 1)
    return gx, (cx * gdim[1] + coord(p[:, 1], lo[1], inv, gdim[1])) * dimz + coord(p[:, 2], lo[2], invz, dimz)

def exchange(arr_left, arr_right):
    """one message to and from each neighbour (fp_shard.cu: an ncclSend/ncclRecv pair per face)"""
    out = {}
    for q, buf in ((rank - 1, arr_left), (rank + 1, arr_right)):
        if 0 <= q < world:
            t = torch.from_numpy(np.ascontiguousarray(buf))
            shp = torch.tensor(list(t.shape) + [0] * (2 - t.dim()), dtype=torch.int64)
            ops = [dist.P2POp(dist.isend, shp, q)]
            rshp = torch.zeros(2, dtype=torch.int64)
            ops.append(dist.P2POp(dist.irecv, rshp, q))
            for w in dist.batch_isend_irecv(ops): w.wait()
            shape = [int(x) for x in rshp if int(x)] or [0]
            r = torch.zeros(shape, dtype=t.dtype)
            ops = [dist.P2POp(dist.isend, t, q), dist.P2POp(dist.irecv, r, q)]
            for w in dist.batch_isend_irecv(ops): w.wait()
            out[q] = r.numpy()
    return out

def binning(owned_idx, p):
    """slab_rebin: sort the owned records, exchange boundary layers' cell counts and records,
    lay out [ghost L | owned | ghost R] and the cell table."""
    gx, key = keys_of(p[owned_idx])
    assert ((gx >= xs0) & (gx < xs1)).all()
    order = np.argsort(key, kind="stable")
    own, key = owned_idx[order], key[order]
    cnt = np.bincount(key, minlength=ldim0 * CL).astype(np.int64)
    assert cnt[:CL].sum() == 0 and cnt[(ldim0 - 1) * CL:].sum() == 0      # ghost layers hold no owned record
    first = own[key < 2 * CL]                                            # my first / last owned layer, verbatim
    last = own[key >= (ldim0 - 2) * CL]
    got_c = exchange(cnt[CL:2 * CL], cnt[(ldim0 - 2) * CL:(ldim0 - 1) * CL])
    got_r = exchange(first, last)
    ghostL = got_r.get(rank - 1, np.zeros(0, np.int64))
    ghostR = got_r.get(rank + 1, np.zeros(0, np.int64))
    if rank - 1 in got_c: cnt[:CL] = got_c[rank - 1]
    if rank + 1 in got_c: cnt[(ldim0 - 1) * CL:] = got_c[rank + 1]
    cell_start = np.concatenate([[0], np.cumsum(cnt)])
    records = np.concatenate([ghostL, own, ghostR])
    assert cell_start[CL] == len(ghostL) and cell_start[-1] == len(records)
    return records, key, cell_start, len(ghostL), len(own)

def walk(records, key, cell_start, nL, nO, p):
    """neighbour counts of the owned slots from the 27 cells (9 rows of z slices) around the HOME cell"""
    out = np.zeros(nO, np.int64)
    P = p[records].astype(np.float64)
    for k in range(nO):
        s = nL + k
        cz = key[k] % dimz; t = key[k] // dimz; cy = t % gdim[1]; cx = t // gdim[1]
        z0, z1 = max(cz - ZSPAN, 0), min(cz + ZSPAN, dimz - 1)
        c = 0
        for x in range(max(cx - 1, 0), min(cx + 1, ldim0 - 1) + 1):
            for y in range(max(cy - 1, 0), min(cy + 1, gdim[1] - 1) + 1):
                rb = (x * gdim[1] + y) * dimz
                a, b = cell_start[rb + z0], cell_start[rb + z1 + 1]
                if b > a:
                    d = P[a:b] - P[s]
                    c += int(((d * d).sum(axis=1) < REACH * REACH).sum())
        out[k] = c - 1                                                    # minus itself
    return out

gxg = coord(pos[:, 0], lo[0], inv, gdim[0])
owned = np.nonzero((gxg >= xs0) & (gxg < xs1))[0]
records, key, cell_start, nL, nO = binning(owned, pos)
truth = np.array([len(v) - 1 for v in cKDTree(pos.astype(np.float64)).query_ball_point(pos.astype(np.float64), REACH - 1e-9)])
got0 = walk(records, key, cell_start, nL, nO, pos)
assert np.array_equal(got0, truth[records[nL:nL + nO]]), "neighbour counts right after a binning"
# lazy steps: everybody drifts by up to skin / 2; the owners' boundary layers are "pushed" into the
# neighbours' ghost blocks (here: the ghost records simply follow the global array) -- the standing
# binning still sees every neighbour
rng = np.random.default_rng(5)
d = rng.normal(size=pos.shape); d /= np.linalg.norm(d, axis=1, keepdims=True)
moved = (pos.astype(np.float64) + d * (SKIN / 2) * rng.random((n, 1))).astype(f32)
got1 = walk(records, key, cell_start, nL, nO, moved)
truth1 = np.array([len(v) - 1 for v in cKDTree(moved.astype(np.float64)).query_ball_point(moved.astype(np.float64), REACH - 1e-9)])
assert np.array_equal(got1, truth1[records[nL:nL + nO]]), "neighbour counts on the standing binning after drift"
# every boid owned exactly once
tot = torch.tensor([nO]); dist.all_reduce(tot); assert int(tot) == n
dist.barrier(); dist.destroy_process_group()
print("SLAB_OK", rank, nL, nO, len(records) - nL - nO)
'''


def test_gloo_world3_slab_protocol(tmp_path):
    """Three CPU ranks run the slab layout protocol of fp_shard.cu (contiguous ghost layers under
    x-slowest keys, boundary layers and their cell counts copied verbatim, walk from home cells)
    and must see exactly the global neighbour counts, right after a binning and after drift."""
    script = tmp_path / "slab_worker.py"
    script.write_text(SLAB_WORKER)
    env = dict(os.environ, FP_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=3",
           "--master-addr", "127.0.0.1", "--master-port", "29521", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and r.stdout.count("SLAB_OK") == 3, r.stdout[-3000:] + r.stderr[-3000:]
