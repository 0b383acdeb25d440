"""Chunk-range sharding of a compressed batch across the GPUs of one box (one process per GPU).

Chunks are fully independent (reference examples/low_level_quickstart_example.cpp:106-108: "chunks can be
re-arranged as well as decompressed with other chunks"), so the hot path shards with NO data-path
collective: rank r owns a contiguous chunk range.  The only exchange is the one-off distribution of the
compressed slab + its (offset, size) table from the rank that holds it -- a broadcast (as BASELINE.json's
north_star states) or a scatter of just each rank's slice -- over NCCL (NVLink 5 / NVSwitch).  The
reference's only multi-GPU precedent moved compressed chunks with peer cudaMemcpyAsync
(benchmarks/benchmark_allgather.cpp:157-197); this is its one-process-per-GPU equivalent.

Works with any torch.distributed backend (nccl on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def partition_equal(n_chunks: int, world: int) -> list[tuple[int, int]]:
    """rank r owns [r*n/R, (r+1)*n/R)  (SURVEY.md 8e)."""
    return [(n_chunks * r // world, n_chunks * (r + 1) // world) for r in range(world)]


def partition_by_bytes(comp_sizes: np.ndarray, world: int) -> list[tuple[int, int]]:
    """Contiguous ranges balanced by the sum of compressed bytes (chunk sizes may vary)."""
    sizes = np.asarray(comp_sizes, dtype=np.int64)
    n = len(sizes)
    if n == 0:
        return [(0, 0)] * world
    csum = np.cumsum(sizes)
    total = int(csum[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        b = int(np.searchsorted(csum, target, side="left")) + 1
        b = min(max(b, bounds[-1]), n)
        bounds.append(b)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def broadcast_batch(slab: torch.Tensor | None, offsets: np.ndarray | None, sizes: np.ndarray | None,
                    src: int = 0, device: torch.device | str = "cpu"):
    """Broadcast a compressed slab and its (offset, size) table from `src` to every rank.
    Returns (slab, offsets, sizes) on every rank."""
    rank = dist.get_rank()
    meta = torch.zeros(2, dtype=torch.int64, device=device)
    if rank == src:
        meta[0] = slab.numel()
        meta[1] = len(sizes)
    dist.broadcast(meta, src)
    nbytes, n = int(meta[0].item()), int(meta[1].item())
    table = torch.zeros(2 * n, dtype=torch.int64, device=device)
    if rank == src:
        table[:n] = torch.from_numpy(np.asarray(offsets, dtype=np.int64)).to(device)
        table[n:] = torch.from_numpy(np.asarray(sizes, dtype=np.int64)).to(device)
    else:
        slab = torch.empty(nbytes, dtype=torch.uint8, device=device)
    dist.broadcast(table, src)
    dist.broadcast(slab, src)
    t = table.cpu().numpy()
    return slab, t[:n].copy(), t[n:].copy()


def scatter_batch(slab: torch.Tensor | None, offsets: np.ndarray | None, sizes: np.ndarray | None,
                  src: int = 0, device: torch.device | str = "cpu"):
    """Send every rank only its own contiguous slice (R x less traffic than a broadcast).
    Returns (local_slab, local_offsets, local_sizes, (begin, end))."""
    rank, world = dist.get_rank(), dist.get_world_size()
    meta = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == src:
        meta[0] = len(sizes)
    dist.broadcast(meta, src)
    n = int(meta[0].item())
    table = torch.zeros(2 * n, dtype=torch.int64, device=device)
    if rank == src:
        table[:n] = torch.from_numpy(np.asarray(offsets, dtype=np.int64)).to(device)
        table[n:] = torch.from_numpy(np.asarray(sizes, dtype=np.int64)).to(device)
    dist.broadcast(table, src)
    t = table.cpu().numpy()
    offs, szs = t[:n], t[n:]
    ranges = partition_by_bytes(szs, world)

    def span(r):
        b, e = ranges[r]
        if e <= b:
            return 0, 0
        return int(offs[b]), int(offs[e - 1] + szs[e - 1])

    lo, hi = span(rank)
    local = torch.empty(max(hi - lo, 1), dtype=torch.uint8, device=device)
    if rank == src:
        reqs = []
        for r in range(world):
            a, z = span(r)
            if r == src:
                local[: z - a] = slab[a:z]
            elif z > a:
                reqs.append(dist.isend(slab[a:z].contiguous(), r))
        for q in reqs:
            q.wait()
    elif hi > lo:
        dist.recv(local[: hi - lo], src)
    b, e = ranges[rank]
    return local, offs[b:e] - lo, szs[b:e].copy(), (b, e)


def exchange_demo(dense: torch.Tensor, c_offs: np.ndarray, c_sizes: np.ndarray, rank: int, world: int) -> dict:
    """bench.py helper: time the NCCL distribution of rank 0's compressed slab (broadcast, as north_star
    words it) with CUDA events, max over ranks, and verify every rank received identical bytes."""
    dev = dense.device
    nbytes = int(c_offs[-1] + c_sizes[-1])
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    slab, offs, sizes = broadcast_batch(dense[:nbytes] if rank == 0 else None, c_offs, c_sizes, 0, dev)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    chk = slab.to(torch.int64).sum().reshape(1)
    allchk = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allchk, chk)
    same = all(int(c.item()) == int(allchk[0].item()) for c in allchk)
    ranges = partition_equal(len(sizes), world)
    return {"what": "ncclBroadcast of rank 0's compressed slab + (offset,size) table; rank r then owns chunk range "
                    f"{ranges[rank] if rank == 0 else ''} ... (equal split)",
            "bytes": int(slab.numel()), "ms": round(float(ms.item()), 3),
            "GBps": round(slab.numel() / (float(ms.item()) * 1e-3) / 1e9, 1), "identical_on_all_ranks": same}
