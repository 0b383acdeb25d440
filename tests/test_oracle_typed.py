"""CPU: the oracle's Cascaded / Bitcomp / ANS restatements round-trip (stream parity is unpinned for these
formats -- the reference bitstreams are undocumented -- so lossless round trip is the property)."""
import numpy as np
import pytest

from conftest import sample_inputs

INPUTS = sample_inputs()
TYPES = {0: 1, 1: 1, 2: 2, 3: 2, 4: 4, 5: 4, 6: 8, 7: 8}


def typed_inputs(ts):
    return {k: v[: len(v) // ts * ts] for k, v in INPUTS.items()}


@pytest.mark.parametrize("type_id", sorted(TYPES))
@pytest.mark.parametrize("layers", [(0, 0, 1), (1, 0, 1), (1, 1, 1), (2, 1, 1), (2, 2, 0), (0, 1, 1), (3, 2, 1)])
def test_cascaded_oracle_roundtrip(oracle, type_id, layers):
    r, d, bp = layers
    for name, data in typed_inputs(8).items():
        comp = oracle.compress_typed("cascaded", data, type=type_id, num_RLEs=r, num_deltas=d, use_bp=bp)
        assert oracle.size("cascaded", comp) == len(data)
        assert oracle.decompress("cascaded", comp, len(data)) == data, name


@pytest.mark.parametrize("type_id", sorted(TYPES))
@pytest.mark.parametrize("algo", [0, 1])
def test_bitcomp_oracle_roundtrip(oracle, type_id, algo):
    for name, data in typed_inputs(8).items():
        comp = oracle.compress_typed("bitcomp", data, algo=algo, type=type_id)
        assert oracle.size("bitcomp", comp) == len(data)
        assert oracle.decompress("bitcomp", comp, len(data)) == data, name


def test_ans_oracle_roundtrip(oracle):
    for name, data in INPUTS.items():
        comp = oracle.compress_typed("ans", data)
        assert oracle.size("ans", comp) == len(data)
        assert oracle.decompress("ans", comp, len(data)) == data, name
    # entropy coder actually compresses low-entropy bytes to ~2 bits/byte
    comp = oracle.compress_typed("ans", INPUTS["gen_data3"])
    assert len(comp) < 0.27 * len(INPUTS["gen_data3"])


def test_typed_oracle_rejects_garbage(oracle):
    for codec in ("cascaded", "bitcomp", "ans"):
        assert oracle.decompress(codec, b"", 100) is None
        assert oracle.decompress(codec, bytes(64), 100) is None
    comp = oracle.compress_typed("ans", INPUTS["lowentropy"])
    assert oracle.decompress("ans", comp[:-8], 65536) is None
    comp = oracle.compress_typed("cascaded", INPUTS["sorted_i64"], type=6, num_RLEs=1, num_deltas=1, use_bp=1)
    assert oracle.decompress("cascaded", comp[: len(comp) // 2], 65536) is None
    assert oracle.decompress("cascaded", comp, 1000) is None


@pytest.mark.parametrize("kind", ["cascaded", "bitcomp"])
def test_oracle_ragged_lengths(oracle, kind):
    """Chunk lengths that are not a multiple of the element size keep their trailing bytes."""
    base = INPUTS["sorted_i64"]
    for type_id in (2, 4, 6):
        kw = dict(type=type_id, num_RLEs=1, num_deltas=1, use_bp=1) if kind == "cascaded" else dict(algo=1, type=type_id)
        for n in (0, 1, 3, 7, 9, 1001, 4099, 65533):
            comp = oracle.compress_typed(kind, base[:n], **kw)
            assert oracle.size(kind, comp) == n
            assert oracle.decompress(kind, comp, n) == base[:n], (kind, type_id, n)
