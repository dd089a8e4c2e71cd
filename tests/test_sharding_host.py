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
