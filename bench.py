#!/usr/bin/env python3
"""bench.py -- the contract benchmark of the batched-decompress hot path.

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--codec snappy|lz4|cascaded|bitcomp|ans]

One "step" = one nvcompBatched<Fmt>DecompressAsync pass (through the C ABI of libnvcomp.so)
over one batch of 10,000 x 64 KB synthetic chunks per GPU.  Default workload = BASELINE.json
configs[1]: Snappy batched decompress, 10,000 x 64 KB synthetic tabular float32 chunks, 1 GPU.
The metric is the reference's: total uncompressed bytes / (1e9 * seconds), CUDA-event timed around the
async call (reference benchmarks/benchmark_template_chunked.cuh:519-539,604-607).

Prints ONE JSON line (rank 0).  `value` = device-resident throughput; `e2e` = the same work with
HOST buffers (pinned): H2D of the compressed chunks + decompress + D2H of the decompressed chunks,
all inside the timed region; `roofline` = algorithmic bytes / event-timed launch duration vs the
measured HBM peak; `cpu_baseline` = the CPU oracle port on this box's host cores (rank 0, N=1).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHUNK = 65536
CHUNKS_PER_GPU = 10000
FMT = {"lz4": "LZ4", "snappy": "Snappy", "cascaded": "Cascaded", "bitcomp": "Bitcomp", "ans": "ANS"}
ORACLE_ID = {"lz4": 0, "snappy": 1, "cascaded": 2, "bitcomp": 3, "ans": 4}
DEFAULT_DATASET = {"lz4": "lz4_mixed", "snappy": "tabular_f32", "cascaded": "sorted_i64",
                   "bitcomp": "sorted_i64", "ans": "lowentropy_bytes"}
KERNEL_NAME = {"lz4": "lz4_decompress_v2_kernel<10>", "snappy": "snappy_decompress_v2_kernel<10>",
               "cascaded": "cascaded_decompress_kernel", "bitcomp": "bitcomp_decompress_kernel",
               "ans": "ans_decompress_kernel"}
WORKLOAD_NAME = {
    "snappy": "BASELINE configs[1]: Snappy batched decompress, 10000x64KB synthetic tabular float32 chunks per GPU",
    "lz4": "BASELINE configs[4] per-GPU share: LZ4 batched decompress, 10000x64KB chunks (run-length int32 + tabular float32) per GPU",
    "cascaded": "BASELINE configs[2]: Cascaded (RLE+delta+bitpack) decompress, sorted int64, 10000x64KB per GPU",
    "bitcomp": "Bitcomp decompress, sorted int64, 10000x64KB per GPU",
    "ans": "BASELINE configs[3]: ANS batched decompress, 10000x64KB low-entropy byte chunks per GPU",
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def codec_opts(kind: str, dataset: str):
    from nvcomp_b200._lib import BitcompOpts, CascadedOpts, Type
    if kind == "cascaded":
        return CascadedOpts(4096, Type.LONGLONG if "i64" in dataset else Type.INT, 1, 1, 1)
    if kind == "bitcomp":
        return BitcompOpts(0, Type.ULONGLONG if "i64" in dataset else Type.UINT)
    return None


# ----------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# CPU oracle (baseline legs only)
# ----------------------------------------------------------------------------------------------
def load_oracle():
    path = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-C", ROOT, "oracle/liboracle.so"], check=True, stdout=subprocess.DEVNULL)
    lib = C.CDLL(path)
    lib.oracle_batch_decompress.restype = C.c_double
    lib.oracle_batch_decompress.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                            C.c_size_t, C.c_void_p, C.c_int]
    return lib


def cpu_decode_time(lib, kind, comp_host: np.ndarray, offs: np.ndarray, lens: np.ndarray, out_host: np.ndarray,
                    threads: int) -> float:
    out_len = np.zeros(len(offs), dtype=np.uint64)
    t = lib.oracle_batch_decompress(ORACLE_ID[kind], comp_host.ctypes.data, offs.ctypes.data, lens.ctypes.data,
                                    len(offs), out_host.ctypes.data, CHUNK, out_len.ctypes.data, threads)
    if t < 0 or not (out_len == CHUNK).all():
        raise RuntimeError("CPU oracle failed to decode the sample")
    return t


def cpu_baseline(kind, comp_host, offs, lens, raw_check: np.ndarray | None, budget_s: float = 4.0):
    lib = load_oracle()
    threads = os.cpu_count() or 1
    n = len(offs)
    out_host = np.empty(n * CHUNK, dtype=np.uint8)
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    lens = np.ascontiguousarray(lens, dtype=np.uint64)
    t0 = cpu_decode_time(lib, kind, comp_host, offs, lens, out_host, threads)   # warm-up + correctness
    if raw_check is not None and not np.array_equal(out_host[: raw_check.size], raw_check.reshape(-1)):
        raise RuntimeError("CPU oracle output differs from the original data")
    reps, times = 0, []
    t_start = time.time()
    while reps < 3 or (time.time() - t_start < budget_s and reps < 200):
        times.append(cpu_decode_time(lib, kind, comp_host, offs, lens, out_host, threads))
        reps += 1
    best = min(times)
    return {"value": n * CHUNK / best / 1e9, "unit": "GB/s", "cores": threads, "kind": "port",
            "sample": f"{n} chunks x 64 KB of this workload, {reps} repetitions, best wall time, "
                      f"oracle/ C decoder with {threads} pthreads (contiguous chunk range per thread)",
            "median_GBps": n * CHUNK / float(np.median(times)) / 1e9}


# ----------------------------------------------------------------------------------------------
def build_workload(kind: str, dataset: str, n_chunks: int, seed_offset: int):
    """Host data + device batch + compressed batch (compressed on the GPU through the C ABI)."""
    import torch
    from nvcomp_b200 import datagen
    from nvcomp_b200.batched import Batch, Codec
    gen = datagen.DATASETS[dataset]
    try:
        data = gen(n_chunks, seed=1000 * seed_offset + gen.__defaults__[0]) if seed_offset else gen(n_chunks)
    except TypeError:
        data = gen(n_chunks)
    slab = torch.from_numpy(data.reshape(-1)).cuda()
    offsets = np.arange(n_chunks, dtype=np.int64) * CHUNK
    inp = Batch(slab, torch.from_numpy(offsets + slab.data_ptr()).cuda(),
                torch.full((n_chunks,), CHUNK, dtype=torch.int64, device="cuda"), offsets)
    codec = Codec(FMT[kind], opts=codec_opts(kind, dataset))
    comp = codec.compress(inp, max_chunk=CHUNK)
    torch.cuda.synchronize()
    return data, inp, codec, comp


def compact(comp, align=16):
    """Pack the compressed chunks contiguously (what a file / network sender would hold)."""
    import torch
    sizes = comp.sizes.cpu().numpy().astype(np.int64)
    al = (sizes + align - 1) // align * align
    offs = np.concatenate([[0], np.cumsum(al)[:-1]]).astype(np.int64)
    total = int(al.sum())
    dense = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
    for i in range(len(sizes)):
        o, src_o, sz = int(offs[i]), int(comp.offsets[i]), int(sizes[i])
        dense[o: o + sz] = comp.slab[src_o: src_o + sz]
    return dense, offs, sizes


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from nvcomp_b200.batched import Batch, empty_batch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", torch.cuda.current_device())

    kind = args.codec
    dataset = args.dataset or DEFAULT_DATASET[kind]
    n = args.chunks
    data, inp, codec, comp_strided = build_workload(kind, dataset, n, rank)
    dense, c_offs, c_sizes = compact(comp_strided)
    del comp_strided
    comp = Batch(dense, torch.from_numpy(c_offs + dense.data_ptr()).cuda(), torch.from_numpy(c_sizes).cuda(), c_offs)
    comp_total = int(c_sizes.sum())
    total = n * CHUNK

    # ---- optional NCCL distribution step (north_star: rank 0 broadcasts the compressed slab + the
    # (offset, size) table, every rank decodes its own chunk range).  Measured, not part of `value`.
    distribute = None
    if world > 1:
        from nvcomp_b200 import shard
        distribute = shard.exchange_demo(dense, c_offs, c_sizes, rank, world)

    out = empty_batch(n, CHUNK)
    caps = inp.sizes
    tb = codec.decompress_get_temp_size(n, CHUNK)
    temp = torch.empty(max(tb, 1), dtype=torch.uint8, device=dev)
    actual = torch.zeros(n, dtype=torch.int64, device=dev)
    status = torch.full((n,), -1, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream

    def step():
        codec.decompress_async(comp.ptrs.data_ptr(), comp.sizes.data_ptr(), caps.data_ptr(), actual.data_ptr(), n,
                               temp.data_ptr(), tb, out.ptrs.data_ptr(), status.data_ptr(), sh)

    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()          # sampled from the warm-up on: same kernel, same load as the timed region
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    # parity gate before timing: bit-exact vs the original data, every status success
    assert bool((status == 0).all().item()) and bool((actual == CHUNK).all().item()), "decompress status/size"
    assert torch.equal(out.slab[:total], inp.slab[:total]), "decompressed bytes differ from the input"

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for a, b in evs:
        a.record()
        step()
        b.record()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    elapsed_ms = e0.elapsed_time(e1)
    launch_ms = [a.elapsed_time(b) for a, b in evs]
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        cs = torch.tensor([float(comp_total)], device=dev)
        dist.all_reduce(cs, op=dist.ReduceOp.SUM)
        comp_total_all = float(cs.item())
    else:
        comp_total_all = float(comp_total)
    if rank == 0:
        # the timed region lasts only tens of ms; keep the identical load running (untimed) until the
        # sampler has seen ~0.6 s of it, so the clock record describes this kernel under load
        t_load = time.time()
        while time.time() - t_load < 0.6:
            for _ in range(10):
                step()
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed region + 0.6 s of the same launches (untimed), nvidia-smi -lms 20"

    # ---- e2e: host buffers in, host buffers out, through the same C-ABI call, pipelined in slices
    e2e = run_e2e(codec, comp, c_offs, c_sizes, inp, n, args, world)
    e2e_dev = run_e2e(codec, comp, c_offs, c_sizes, inp, n, args, world, copy_back=False)

    ms_per_step = elapsed_ms / args.steps
    value = world * total / (ms_per_step * 1e-3) / 1e9
    peak, peak_src = peaks()
    avg_launch_ms = float(np.mean(launch_ms))
    alg_bytes = total + comp_total + 44 * n
    achieved = alg_bytes / (avg_launch_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(f"{kind}:{dataset}")

    line = {
        "metric": "decompressed GB/s (64KB chunks), whole job",
        "value": round(value, 2), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME[kind], "codec": kind, "dataset": dataset, "chunks_per_gpu": n,
                   "chunk_bytes": CHUNK, "compression_ratio": round(world * total / comp_total_all, 3),
                   "l2_policy": "inputs larger than L2 (compressed + decompressed footprint per step = "
                                f"{(total + comp_total) / 1e6:.0f} MB vs 126 MB L2)",
                   "sharding": "contiguous chunk range per rank, no data-path collective" if world > 1 else "single GPU"},
        "e2e": e2e,
        "e2e_device_consumer": e2e_dev,
        "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "kernel": KERNEL_NAME[kind], "avg_launch_ms": round(avg_launch_ms, 4)},
        "clocks": clocks,
    }
    if distribute is not None:
        line["distribute"] = distribute
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = min(n, args.cpu_chunks)
        host = dense[: int(c_offs[sample - 1] + c_sizes[sample - 1])].cpu().numpy() if sample else np.zeros(1, np.uint8)
        line["cpu_baseline"] = cpu_baseline(kind, host, c_offs[:sample], c_sizes[:sample], data[:sample])
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(codec, comp, c_offs, c_sizes, inp, n, args, world, copy_back=True):
    """Same decode through the C ABI, but the compressed chunks start in pinned HOST memory and the
    decompressed chunks end in pinned HOST memory; both copies are inside the timed region.  The batch is
    processed in slices on three streams so H2D, decode and D2H overlap."""
    import torch
    import torch.distributed as dist
    from nvcomp_b200.batched import empty_batch
    dev = comp.slab.device
    total = n * CHUNK
    comp_total = int(c_offs[-1] + c_sizes[-1])
    h_comp = torch.empty(comp_total, dtype=torch.uint8).pin_memory()
    h_comp.copy_(comp.slab[:comp_total])
    h_out = torch.empty(total, dtype=torch.uint8).pin_memory()
    d_comp = torch.empty(comp_total + 64, dtype=torch.uint8, device=dev)
    out = empty_batch(n, CHUNK)
    ptrs = torch.from_numpy(c_offs + d_comp.data_ptr()).to(dev)
    sizes = torch.from_numpy(c_sizes).to(dev)
    caps = inp.sizes
    actual = torch.zeros(n, dtype=torch.int64, device=dev)
    status = torch.full((n,), -1, dtype=torch.int32, device=dev)
    h_status = torch.empty(n, dtype=torch.int32).pin_memory()
    nslices = 4
    bounds = [n * i // nslices for i in range(nslices + 1)]
    # one stream per slice: H2D -> decode -> D2H in order on that stream; the slices' copies share the
    # two copy engines and their decode kernels overlap (a quarter batch does not fill the GPU)
    streams = [torch.cuda.Stream() for _ in range(nslices)]
    tb = codec.decompress_get_temp_size(n, CHUNK)
    temps = [torch.empty(max(tb, 1), dtype=torch.uint8, device=dev) for _ in range(nslices)]

    def step():
        for i in range(nslices):
            a, b = bounds[i], bounds[i + 1]
            lo, hi = int(c_offs[a]), int(c_offs[b - 1] + c_sizes[b - 1])
            st = streams[i]
            with torch.cuda.stream(st):
                d_comp[lo:hi].copy_(h_comp[lo:hi], non_blocking=True)
                codec.decompress_async(ptrs.data_ptr() + 8 * a, sizes.data_ptr() + 8 * a, caps.data_ptr() + 8 * a,
                                       actual.data_ptr() + 8 * a, b - a, temps[i].data_ptr(), tb,
                                       out.ptrs.data_ptr() + 8 * a, status.data_ptr() + 4 * a, st.cuda_stream)
                if copy_back:
                    h_out[a * CHUNK: b * CHUNK].copy_(out.slab[a * CHUNK: b * CHUNK], non_blocking=True)
                h_status[a:b].copy_(status[a:b], non_blocking=True)

    def drain():
        for st in streams:
            st.synchronize()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    ok = bool((h_status == 0).all().item()) and (not copy_back or np.array_equal(
        h_out.numpy()[: 4 * CHUNK], inp.slab[: 4 * CHUNK].cpu().numpy()))
    assert ok, "e2e output mismatch"
    if world > 1:
        dist.barrier()
    steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
        drain()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if not copy_back:
        return {"value": round(world * total / (ms * 1e-3) / 1e9, 2), "unit": "GB/s",
                "h2d_bytes_per_step": comp_total, "d2h_bytes_per_step": 4 * n, "ms_per_step": round(ms, 3),
                "what": "same call, decompressed chunks stay in HBM for a GPU consumer; only statuses return"}
    return {"value": round(world * total / (ms * 1e-3) / 1e9, 2), "unit": "GB/s",
            "h2d_bytes_per_step": comp_total, "d2h_bytes_per_step": total + 4 * n,
            "ms_per_step": round(ms, 3), "pipeline": f"{nslices} slices, one stream each (H2D -> decode -> D2H)",
            "what": "pinned host compressed chunks -> H2D -> nvcompBatched*DecompressAsync -> D2H of the "
                    "decompressed chunks and statuses into pinned host memory (PCIe-bound: "
                    "d2h_bytes/ms_per_step is the link rate)"}


def oracle_compress_batch(lib, kind, dataset, data: np.ndarray):
    """Compress every chunk with the CPU oracle's own encoders (reference arm: no CUDA code of this repo runs)."""
    from concurrent.futures import ThreadPoolExecutor
    u8p, sz = C.c_char_p, C.c_size_t
    n = data.shape[0]
    cap = 2 * CHUNK + 65536
    t64 = "i64" in dataset
    if kind in ("lz4", "snappy", "ans"):
        fn = getattr(lib, f"oracle_{kind}_compress")
        fn.argtypes, fn.restype = [u8p, sz, u8p, sz], C.c_long
        call = lambda raw, out: fn(raw, len(raw), out, cap)
    elif kind == "cascaded":
        fn = lib.oracle_cascaded_compress
        fn.argtypes, fn.restype = [u8p, sz, u8p, sz, sz, C.c_uint, C.c_int, C.c_int, C.c_int], C.c_long
        call = lambda raw, out: fn(raw, len(raw), out, cap, 4096, 6 if t64 else 4, 1, 1, 1)
    else:
        fn = lib.oracle_bitcomp_compress
        fn.argtypes, fn.restype = [u8p, sz, u8p, sz, C.c_uint, C.c_uint], C.c_long
        call = lambda raw, out: fn(raw, len(raw), out, cap, 0, 7 if t64 else 5)

    def one(i):
        out = C.create_string_buffer(cap)
        r = call(data[i].tobytes(), out)
        assert r > 0
        return out.raw[:r]

    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
        comps = list(ex.map(one, range(n)))
    sizes = np.array([len(c) for c in comps], dtype=np.int64)
    al = (sizes + 15) // 16 * 16
    offs = np.concatenate([[0], np.cumsum(al)[:-1]]).astype(np.int64)
    slab = np.zeros(int(al.sum()) + 64, dtype=np.uint8)
    for c, o in zip(comps, offs):
        slab[o:o + len(c)] = np.frombuffer(c, dtype=np.uint8)
    return slab, offs, sizes


def run_reference(args):
    """--impl reference: the reference's own implementation of this path is the closed libnvcomp.so
    (not in /root/reference, not installable: no source, no wheel).  Per the task's tier rules this arm
    times the CPU implementation of the path instead: the oracle port (oracle/*.c) on all host cores,
    on a bounded sample of the same workload.  Nothing of libnvcomp.so is loaded on this arm: the synthetic
    chunks are compressed by the oracle's own CPU encoders and decoded by its decoders."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    kind = args.codec
    dataset = args.dataset or DEFAULT_DATASET[kind]
    n = min(args.chunks, args.cpu_chunks)
    from nvcomp_b200 import datagen
    data = datagen.DATASETS[dataset](n)
    lib = load_oracle()
    host, c_offs, c_sizes = oracle_compress_batch(lib, kind, dataset, data)
    threads = os.cpu_count() or 1
    offs = np.ascontiguousarray(c_offs, dtype=np.uint64)
    lens = np.ascontiguousarray(c_sizes, dtype=np.uint64)
    out_host = np.empty(n * CHUNK, dtype=np.uint8)
    for _ in range(max(args.warmup, 1)):
        cpu_decode_time(lib, kind, host, offs, lens, out_host, threads)
    assert np.array_equal(out_host, data.reshape(-1)), "oracle output differs from the original data"
    times = [cpu_decode_time(lib, kind, host, offs, lens, out_host, threads) for _ in range(args.steps)]
    sec = float(np.mean(times))
    v = n * CHUNK / sec / 1e9
    sample = (f"{n} chunks x 64 KB per step ({'same batch size' if n == args.chunks else 'bounded sample'} as the GPU arm's "
              f"workload, same generator), oracle/ C encoders + decoders, {threads} pthreads, mean of {args.steps} steps")
    line = {
        "impl": "reference", "metric": "decompressed GB/s (64KB chunks), whole job", "value": round(v, 2),
        "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1),
        "ms_per_step": round(sec * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME[kind], "codec": kind, "dataset": dataset, "chunks_per_step": n,
                   "chunk_bytes": CHUNK, "compression_ratio": round(n * CHUNK / float(c_sizes.sum()), 3),
                   "note": "the reference library is closed-source and absent; CPU implementation of the path "
                           "(oracle port) on the host cores, per the task's reference-arm rule"},
        "cpu_baseline": {"value": round(v, 2), "unit": "GB/s", "cores": threads, "kind": "port", "sample": sample,
                         "best_GBps": round(n * CHUNK / min(times) / 1e9, 2)},
        "e2e": {"value": round(v, 2), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--codec", default="snappy", choices=sorted(FMT))
    ap.add_argument("--dataset", default=None)
    ap.add_argument("--chunks", type=int, default=CHUNKS_PER_GPU)
    ap.add_argument("--cpu-chunks", type=int, default=10000, help="bounded CPU sample (chunks)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
