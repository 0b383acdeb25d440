"""Deterministic synthetic workloads for the tests and bench.py (host side, numpy).

The reference defines only ``gen_data`` and the all-zero / all-random extremes;
the tabular generators below are the frozen definitions of BASELINE.json's
config lines.  Every generator returns a ``uint8`` array of
shape ``(n_chunks, chunk_bytes)``.
"""
from __future__ import annotations

import numpy as np

CHUNK = 65536


def ref_gen_data(max_byte: int, size: int, rs: np.random.RandomState) -> np.ndarray:
    """Bit-exact restatement of the reference's ``gen_data`` (benchmarks/benchmark_common.h:158-175):
    ``std::uniform_int_distribution<uint16_t>(0, max_byte)`` drawn from a caller-owned
    ``std::mt19937`` -- libstdc++'s down-scaling with rejection over 32-bit draws.
    ``rs`` must be ``np.random.RandomState(seed)`` (init_genrand seeding == std::mt19937(seed))."""
    rng_range = max_byte + 1
    scaling = (1 << 32) // rng_range
    past = rng_range * scaling
    out = np.empty(size, dtype=np.uint8)
    filled = 0
    bitgen = rs._bit_generator            # raw 32-bit mt19937 outputs, the sequence std::mt19937 produces
    block = 1 << 22                       # cache-sized blocks: the 64-bit temporaries stay out of DRAM
    while filled < size:
        raw = bitgen.random_raw(min(block, size - filled))
        if past != (1 << 32):
            raw = raw[raw < past]         # rejected draws are consumed and skipped, exactly like the C++ do/while loop
        ok = raw // scaling
        out[filled:filled + len(ok)] = ok.astype(np.uint8)
        filled += len(ok)
    return out


def snappy_synth(n_chunks: int, max_byte: int = 3, seed: int = 0, chunk: int = CHUNK) -> np.ndarray:
    """Input of the reference's benchmark_snappy_synth (benchmarks/benchmark_snappy_synth.cpp:365-369):
    one mt19937(0) stream, ``n_chunks`` consecutive gen_data(max_byte, 65536) draws."""
    rs = np.random.RandomState(seed)
    return ref_gen_data(max_byte, n_chunks * chunk, rs).reshape(n_chunks, chunk)


def runlength_i32(n_chunks: int, seed: int = 0, chunk: int = CHUNK) -> np.ndarray:
    """cfg1: int32 run-length data -- value ~ U[0, 2^31) repeated L ~ U[1, 256] times."""
    rng = np.random.Generator(np.random.MT19937(seed))
    n = n_chunks * chunk // 4
    n_runs = n // 100 + 1024
    lens = rng.integers(1, 257, size=n_runs)
    while lens.sum() < n:
        lens = np.concatenate([lens, rng.integers(1, 257, size=n_runs)])
    vals = rng.integers(0, 1 << 31, size=len(lens), dtype=np.int64).astype(np.int32)
    data = np.repeat(vals, lens)[:n]
    return data.view(np.uint8).reshape(n_chunks, chunk)


def _f32_price_walk(rng, n):
    x = 100.0 + np.cumsum(rng.normal(0.0, 0.05, size=n))
    return np.round(x, 2).astype(np.float32)


def _f32_lowcard(rng, n):
    levels = np.round(np.arange(0, 11) * 0.01, 2).astype(np.float32)   # TPC-H discount: 0.00..0.10
    return levels[rng.integers(0, 11, size=n)]


def _f32_clustered(rng, n):
    # a column sorted/clustered on its value: geometric run lengths (mean 24) of a slowly rising price
    n_runs = n // 8 + 16
    lens = rng.geometric(1.0 / 24.0, size=n_runs)
    vals = np.round(10.0 + np.cumsum(rng.integers(1, 50, size=n_runs)) * 0.01, 2).astype(np.float32)
    return np.repeat(vals, lens)[:n]


def _f32_sensor(rng, n):
    t = np.arange(n, dtype=np.float64)
    return (2.0 + 0.5 * np.sin(t * 1e-3) + rng.normal(0, 1e-4, size=n)).astype(np.float32)


_F32_COLUMNS = (_f32_price_walk, _f32_lowcard, _f32_clustered, _f32_sensor)
F32_COLUMN_NAMES = ("price_walk", "lowcard", "clustered", "sensor")


def tabular_f32(n_chunks: int, seed: int = 1, chunk: int = CHUNK, column: int | None = None) -> np.ndarray:
    """cfg2(ii): a float32 table stored column-chunked.  Chunk i holds 16384 consecutive
    values of column ``i % 4`` (or of ``column`` when given):
      0 price_walk : random walk rounded to 0.01 (short 4-byte matches)
      1 lowcard    : 11 distinct values (TPC-H l_discount)
      2 clustered  : sorted/clustered column, geometric runs (mean 24 values)
      3 sensor     : smooth series with noise (mantissa nearly incompressible)."""
    rng = np.random.Generator(np.random.MT19937(seed))
    per = chunk // 4
    out = np.empty((n_chunks, per), dtype=np.float32)
    cols = range(4) if column is None else (column,)
    for c in cols:
        idx = np.arange(n_chunks) if column is not None else np.arange(c, n_chunks, 4)
        if len(idx) == 0:
            continue
        vals = _F32_COLUMNS[c](rng, len(idx) * per)
        out[idx] = vals.reshape(len(idx), per)
    return out.view(np.uint8).reshape(n_chunks, chunk)


def sorted_i64(n_chunks: int, seed: int = 2, chunk: int = CHUNK) -> np.ndarray:
    """cfg3: sorted int64 column, v[0] ~ U[0, 2^40), steps geometric(p=0.3)-1 (~30 % duplicates)."""
    rng = np.random.Generator(np.random.MT19937(seed))
    n = n_chunks * chunk // 8
    steps = rng.geometric(0.3, size=n).astype(np.int64) - 1
    v0 = int(rng.integers(0, 1 << 40))
    data = v0 + np.cumsum(steps)
    return data.astype(np.int64).view(np.uint8).reshape(n_chunks, chunk)


def lowentropy_bytes(n_chunks: int, seed: int = 3, chunk: int = CHUNK) -> np.ndarray:
    """cfg4: bytes ~ geometric over 256 symbols, about 2 bits/byte of entropy."""
    rng = np.random.Generator(np.random.MT19937(seed))
    g = rng.geometric(0.5, size=n_chunks * chunk) - 1
    return np.minimum(g, 255).astype(np.uint8).reshape(n_chunks, chunk)


def random_bytes(n_chunks: int, seed: int = 4, chunk: int = CHUNK) -> np.ndarray:
    rng = np.random.Generator(np.random.MT19937(seed))
    return rng.integers(0, 256, size=(n_chunks, chunk), dtype=np.uint8)


def zeros(n_chunks: int, chunk: int = CHUNK) -> np.ndarray:
    return np.zeros((n_chunks, chunk), dtype=np.uint8)


def lz4_mixed(n_chunks: int, seed: int = 5, chunk: int = CHUNK) -> np.ndarray:
    """cfg5: half run-length int32 (cfg1 style), half tabular float32 (cfg2 style), interleaved."""
    a = runlength_i32((n_chunks + 1) // 2, seed=seed, chunk=chunk)
    b = tabular_f32(n_chunks // 2, seed=seed + 1, chunk=chunk)
    out = np.empty((n_chunks, chunk), dtype=np.uint8)
    out[0::2] = a
    out[1::2] = b
    return out


DATASETS = {
    "runlength_i32": runlength_i32,
    "tabular_f32": tabular_f32,
    "sorted_i64": sorted_i64,
    "lowentropy_bytes": lowentropy_bytes,
    "random_bytes": random_bytes,
    "snappy_synth": snappy_synth,
    "lz4_mixed": lz4_mixed,
}
