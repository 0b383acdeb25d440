"""CPU: pin the oracle (oracle/*.c) against the independent codecs the reference links or
names -- liblz4 1.9.4 (reference examples/lz4_cpu_compression.cu, lz4_cpu_decompression.cu)
and pyarrow's bundled snappy -- and against the committed golden vectors."""
import json
import os

import numpy as np
import pytest

from conftest import sample_inputs

INPUTS = sample_inputs()


@pytest.mark.parametrize("name", sorted(INPUTS))
def test_lz4_oracle_decodes_liblz4(oracle, liblz4, name):
    data = INPUTS[name]
    for hc in (0, 12):   # reference uses LZ4_compress_HC level 12 (lz4_cpu_compression.cu:61-66)
        comp = liblz4.compress(data, hc)
        assert oracle.size("lz4", comp) == len(data)
        assert oracle.decompress("lz4", comp, len(data)) == data


@pytest.mark.parametrize("name", sorted(INPUTS))
def test_liblz4_decodes_oracle_lz4(oracle, liblz4, name):
    data = INPUTS[name]
    comp = oracle.compress("lz4", data)
    assert liblz4.decompress(comp, len(data)) == data
    assert oracle.decompress("lz4", comp, len(data)) == data


@pytest.mark.parametrize("name", sorted(INPUTS))
def test_snappy_oracle_vs_pyarrow(oracle, name):
    pa = pytest.importorskip("pyarrow")
    codec = pa.Codec("snappy")
    data = INPUTS[name]
    comp = codec.compress(data).to_pybytes() if len(data) else b"\x00"
    assert oracle.size("snappy", comp) == len(data)
    assert oracle.decompress("snappy", comp, len(data)) == data
    ours = oracle.compress("snappy", data)
    if len(data):
        assert codec.decompress(ours, decompressed_size=len(data)).to_pybytes() == data
    assert oracle.decompress("snappy", ours, len(data)) == data


def test_lz4_oracle_rejects_malformed(oracle):
    good = oracle.compress("lz4", INPUTS["text"])
    n = len(INPUTS["text"])
    assert oracle.decompress("lz4", good[:-3], n) is None          # truncated
    assert oracle.decompress("lz4", good, n - 1) is None           # output too small
    assert oracle.decompress("lz4", b"\x10A\x05\x00", 64) is None  # offset beyond start
    assert oracle.decompress("lz4", b"\x00\x00\x00", 64) is None   # offset 0
    assert oracle.decompress("lz4", b"\xf0", 64) is None           # missing extension byte


def test_snappy_oracle_rejects_malformed(oracle):
    good = oracle.compress("snappy", INPUTS["text"])
    n = len(INPUTS["text"])
    assert oracle.decompress("snappy", good[:-3], n) is None
    assert oracle.decompress("snappy", good, n - 1) is None
    assert oracle.decompress("snappy", b"\x08\x00A\x05\x10", 64) is None   # copy offset beyond start
    assert oracle.decompress("snappy", b"\x05\x00A", 64) is None           # short output
    assert oracle.decompress("snappy", b"\xff\xff\xff\xff\xff\x01", 64) is None  # bad varint


def test_snappy_oracle_all_tag_kinds(oracle):
    # hand-built legal stream exercising literal-with-length-bytes and copy-1 / copy-2 / copy-4
    lit = bytes(range(70))
    stream = bytearray()
    total = 70 + 5 + 20 + 9
    stream += bytes([total])
    stream += bytes([60 << 2, 69]) + lit                 # literal, 1 length byte
    stream += bytes([1 | ((5 - 4) << 2) | (0 << 5), 70])  # copy-1 len 5 off 70
    stream += bytes([2 | ((20 - 1) << 2), 10, 0])         # copy-2 len 20 off 10 (overlapping)
    stream += bytes([3 | ((9 - 1) << 2), 95, 0, 0, 0])    # copy-4 len 9 off 95
    exp = bytearray(lit)
    for ln, off in ((5, 70), (20, 10), (9, 95)):
        for _ in range(ln):
            exp.append(exp[-off])
    assert oracle.decompress("snappy", bytes(stream), total) == bytes(exp)


def test_golden_vectors(oracle, golden_dir):
    """Committed compressed streams produced by liblz4 / pyarrow-snappy (tests/golden/make_golden.py)."""
    man = json.load(open(os.path.join(golden_dir, "manifest.json")))
    assert len(man["vectors"]) >= 10
    for v in man["vectors"]:
        comp = open(os.path.join(golden_dir, v["comp"]), "rb").read()
        raw = open(os.path.join(golden_dir, v["raw"]), "rb").read()
        assert oracle.decompress(v["codec"], comp, len(raw)) == raw, v["comp"]
