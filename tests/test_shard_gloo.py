"""CPU, world_size 2, gloo: the host-side multi-GPU plumbing (chunk-range partition, broadcast / scatter of
the compressed slab + table).  The decode itself is covered by the -m gpu tests."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_partitions_cover_every_chunk_once():
    from nvcomp_b200 import shard
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 1000, 10000):
        sizes = rng.integers(100, 70000, n)
        for world in (1, 2, 3, 8):
            for ranges in (shard.partition_equal(n, world), shard.partition_by_bytes(sizes, world)):
                assert ranges[0][0] == 0 and ranges[-1][1] == n
                for (a, b), (c, d) in zip(ranges, ranges[1:]):
                    assert b == c and a <= b
            if n >= 1000:
                byt = [int(sizes[a:b].sum()) for a, b in shard.partition_by_bytes(sizes, world)]
                assert max(byt) - min(byt) <= 2 * 70000


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nvcomp_b200 import shard
    rng = np.random.default_rng(42)
    sizes = rng.integers(1, 5000, 333).astype(np.int64)
    al = (sizes + 15) // 16 * 16
    offs = np.concatenate([[0], np.cumsum(al)[:-1]]).astype(np.int64)
    slab_np = rng.integers(0, 256, int(al.sum()), dtype=np.uint8)
    slab = torch.from_numpy(slab_np.copy()) if rank == 0 else None
    got, o, s = shard.broadcast_batch(slab, offs if rank == 0 else None, sizes if rank == 0 else None, 0, "cpu")
    ok = np.array_equal(got.numpy(), slab_np) and np.array_equal(o, offs) and np.array_equal(s, sizes)
    local, lo, ls, (b, e) = shard.scatter_batch(slab, offs if rank == 0 else None, sizes if rank == 0 else None, 0, "cpu")
    for i in range(e - b):
        ref = slab_np[offs[b + i]: offs[b + i] + sizes[b + i]]
        ok = ok and np.array_equal(local.numpy()[lo[i]: lo[i] + ls[i]], ref)
    cnt = torch.tensor([e - b])
    dist.all_reduce(cnt)
    ok = ok and int(cnt.item()) == len(sizes)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_and_scatter_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
