"""-m gpu: LZ4 and Snappy batched codecs (CUDA, through the C ABI) against the CPU oracle,
liblz4 / pyarrow-snappy and the committed golden vectors.  Bit-exact."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import sample_inputs

pytestmark = pytest.mark.gpu

INPUTS = sample_inputs()
NAMES = sorted(INPUTS)
FMT = {"lz4": "LZ4", "snappy": "Snappy"}


def _codec(kind, **kw):
    from nvcomp_b200.batched import Codec
    return Codec(FMT[kind], **kw)


def _cpu_compress(kind, oracle, liblz4, data, variant):
    if kind == "lz4":
        if variant == "oracle":
            return oracle.compress("lz4", data)
        return liblz4.compress(data, 12 if variant == "hc" else 0)
    if variant == "oracle":
        return oracle.compress("snappy", data)
    import pyarrow as pa
    return pa.Codec("snappy").compress(data).to_pybytes() if len(data) else b"\x00"


@pytest.mark.parametrize("kind", ["lz4", "snappy"])
@pytest.mark.parametrize("variant", ["lib", "hc", "oracle"])
@pytest.mark.parametrize("misalign", [0, 1, 7])
def test_gpu_decodes_cpu_streams(kind, variant, misalign, oracle, liblz4):
    """CPU-produced streams (liblz4 default / HC-12 as in reference examples/lz4_cpu_compression.cu:61-66,
    pyarrow snappy, oracle encoder) decode bit-exactly on the GPU."""
    from gpu_util import gpu_decompress
    raws = [INPUTS[n] for n in NAMES]
    comps = [_cpu_compress(kind, oracle, liblz4, r, variant) for r in raws]
    outs, actual, status, _ = gpu_decompress(_codec(kind), comps, [len(r) for r in raws], misalign=misalign)
    assert (status == 0).all(), status
    assert actual.tolist() == [len(r) for r in raws]
    for n, o, r in zip(NAMES, outs, raws):
        assert o == r, n


@pytest.mark.parametrize("kind", ["lz4", "snappy"])
@pytest.mark.parametrize("misalign", [0, 3])
def test_cpu_decodes_gpu_streams(kind, misalign, oracle, liblz4):
    """GPU-compressed chunks decode with liblz4's LZ4_decompress_safe (reference
    examples/lz4_cpu_decompression.cu:143-147) / pyarrow snappy and with the oracle."""
    from gpu_util import gpu_compress
    raws = [INPUTS[n] for n in NAMES]
    comps, _ = gpu_compress(_codec(kind), raws, misalign=misalign)
    for n, c, r in zip(NAMES, comps, raws):
        assert oracle.decompress(kind, c, len(r)) == r, n
        if kind == "lz4":
            assert liblz4.decompress(c, len(r)) == r, n
            assert len(c) <= len(r) + len(r) // 255 + 16
        elif len(r):
            import pyarrow as pa
            assert pa.Codec("snappy").decompress(c, decompressed_size=len(r)).to_pybytes() == r, n


@pytest.mark.parametrize("kind", ["lz4", "snappy"])
def test_lz4_type_hints_and_roundtrip(kind, oracle):
    from gpu_util import gpu_compress, gpu_decompress
    from nvcomp_b200._lib import LZ4Opts, Type
    raws = [INPUTS[n] for n in NAMES if len(INPUTS[n]) % 4 == 0]
    opt_list = [None] if kind == "snappy" else [LZ4Opts(Type.CHAR), LZ4Opts(Type.SHORT), LZ4Opts(Type.INT),
                                                LZ4Opts(Type.BITS)]
    for opts in opt_list:
        codec = _codec(kind, opts=opts)
        comps, _ = gpu_compress(codec, raws)
        outs, actual, status, _ = gpu_decompress(codec, comps, [len(r) for r in raws])
        assert (status == 0).all()
        assert outs == raws


@pytest.mark.parametrize("kind", ["lz4", "snappy"])
def test_get_decompress_size(kind, oracle, liblz4):
    from nvcomp_b200.batched import make_batch
    raws = [INPUTS[n] for n in NAMES]
    comps = [_cpu_compress(kind, oracle, liblz4, r, "lib") for r in raws]
    sizes = _codec(kind).get_decompress_size(make_batch(comps))
    torch.cuda.synchronize()
    assert sizes.cpu().tolist() == [len(r) for r in raws]


@pytest.mark.parametrize("kind", ["lz4", "snappy"])
def test_malformed_streams_fail_cleanly(kind, oracle, liblz4):
    """Truncated / corrupt / overflowing chunks: status CannotDecompress, size 0, no fault
    (reference CHANGELOG.md:160-164); good chunks in the same batch still decode."""
    from gpu_util import gpu_decompress
    text = INPUTS["text"]
    good = _cpu_compress(kind, oracle, liblz4, text, "lib")
    bad = [good[:-3], good[: len(good) // 2], good]
    caps = [len(text), len(text), len(text) - 1]
    if kind == "lz4":
        bad += [b"\x10A\x05\x00", b"\x00\x00\x00", b"\xf0", b"\x1fA\x01\x00" + b"\xff" * 40]
    else:
        bad += [b"\x08\x00A\x05\x10", b"\x05\x00A", b"\xff\xff\xff\xff\xff\x01", b"\x40\xfc\xff\xff\xff\xff"]
    caps += [64] * (len(bad) - len(caps))
    chunks = bad + [good]
    caps = caps + [len(text)]
    outs, actual, status, _ = gpu_decompress(_codec(kind), chunks, caps)
    for i in range(len(bad)):
        assert status[i] == 12 and actual[i] == 0, (i, status[i], actual[i])
        assert oracle.decompress(kind, chunks[i], caps[i]) is None, i     # oracle agrees it is malformed
    assert status[-1] == 0 and outs[-1] == text


@pytest.mark.parametrize("kind", ["lz4", "snappy"])
def test_nullable_outputs_and_aliasing(kind, oracle, liblz4):
    """actual_bytes / statuses may be null for LZ4 and Snappy (doc/lowlevel_c_quickstart.md:140);
    actual_bytes may alias the capacity array (benchmarks/benchmark_snappy_synth.cpp:244-245)."""
    from gpu_util import gpu_decompress
    from nvcomp_b200.batched import Codec, empty_batch, make_batch, _stream_handle
    raws = [INPUTS[n] for n in NAMES]
    comps = [_cpu_compress(kind, oracle, liblz4, r, "lib") for r in raws]
    outs, actual, status, _ = gpu_decompress(_codec(kind), comps, [len(r) for r in raws], want_actual=False,
                                             want_status=False)
    assert outs == raws
    codec = _codec(kind)
    comp = make_batch(comps)
    out = empty_batch(len(raws), 65536)
    caps = torch.tensor([len(r) for r in raws], dtype=torch.int64, device="cuda")
    temp = torch.empty(1024, dtype=torch.uint8, device="cuda")
    codec.decompress_async(comp.ptrs.data_ptr(), comp.sizes.data_ptr(), caps.data_ptr(), caps.data_ptr(), len(raws),
                           temp.data_ptr(), 1024, out.ptrs.data_ptr(), None, _stream_handle(None))
    torch.cuda.synchronize()
    assert caps.cpu().tolist() == [len(r) for r in raws]
    assert out.to_host([len(r) for r in raws]) == raws
    # no workspace at all -> static chunk assignment, same result
    out2 = empty_batch(len(raws), 65536)
    codec.decompress_async(comp.ptrs.data_ptr(), comp.sizes.data_ptr(), caps.data_ptr(), None, len(raws),
                           None, 0, out2.ptrs.data_ptr(), None, _stream_handle(None))
    torch.cuda.synchronize()
    assert out2.to_host([len(r) for r in raws]) == raws


def test_golden_vectors_gpu(golden_dir):
    from gpu_util import gpu_decompress
    man = json.load(open(os.path.join(golden_dir, "manifest.json")))
    for kind in ("lz4", "snappy"):
        vs = [v for v in man["vectors"] if v["codec"] == kind]
        comps = [open(os.path.join(golden_dir, v["comp"]), "rb").read() for v in vs]
        raws = [open(os.path.join(golden_dir, v["raw"]), "rb").read() for v in vs]
        outs, actual, status, _ = gpu_decompress(_codec(kind), comps, [len(r) for r in raws])
        assert (status == 0).all()
        assert outs == raws


@pytest.mark.parametrize("kind", ["lz4", "snappy"])
@pytest.mark.parametrize("dataset", ["runlength_i32", "tabular_f32", "snappy_synth", "random_bytes", "lz4_mixed"])
def test_batch_roundtrip_property(kind, dataset):
    """Size-independent property at batch scale: decompress(compress(x)) == x for 2000 x 64 KB chunks,
    compared on the device; sizes and statuses exact."""
    from nvcomp_b200 import datagen
    from nvcomp_b200.batched import empty_batch
    n = 2000
    data = datagen.DATASETS[dataset](n)
    codec = _codec(kind)
    slab = torch.from_numpy(data.reshape(-1)).cuda()
    from nvcomp_b200.batched import Batch
    offsets = np.arange(n, dtype=np.int64) * 65536
    inp = Batch(slab, torch.from_numpy(offsets + slab.data_ptr()).cuda(),
                torch.full((n,), 65536, dtype=torch.int64, device="cuda"), offsets)
    comp = codec.compress(inp, max_chunk=65536)
    out = empty_batch(n, 65536, fill=0x5A)
    actual, status = codec.decompress(comp, out, max_chunk=65536)
    torch.cuda.synchronize()
    assert (status == 0).all().item()
    assert (actual == 65536).all().item()
    assert out.offsets[1] - out.offsets[0] == 65536
    assert torch.equal(out.slab[: n * 65536], slab)
    ratio = n * 65536 / comp.sizes.sum().item()
    assert ratio > 0.9


def test_ragged_and_empty_batch(oracle):
    from gpu_util import gpu_compress, gpu_decompress
    rng = np.random.default_rng(5)
    raws = [bytes(rng.integers(0, 4, int(s), dtype=np.uint8)) for s in [0, 1, 2, 3, 4, 5, 11, 12, 13, 14, 63, 64, 65,
                                                                       4095, 4096, 4097, 65535, 65536, 100000, 262144]]
    for kind in ("lz4", "snappy"):
        codec = _codec(kind)
        comps, _ = gpu_compress(codec, raws)
        for c, r in zip(comps, raws):
            assert oracle.decompress(kind, c, len(r)) == r
        outs, actual, status, _ = gpu_decompress(codec, comps, [len(r) for r in raws])
        assert (status == 0).all() and outs == raws
        # batch of zero chunks is a no-op
        from nvcomp_b200.batched import _stream_handle
        codec.decompress_async(None, None, None, None, 0, None, 0, None, None, _stream_handle(None))


def _lz4_block(seqs, tail_literals):
    """Assemble an LZ4 block from (literals, offset, match_len) sequences + the final literal-only sequence;
    returns (block, expected_output)."""
    def ext(n):
        out = bytearray()
        while n >= 255:
            out.append(255); n -= 255
        out.append(n)
        return bytes(out)
    blk, out = bytearray(), bytearray()
    for lits, off, ml in seqs:
        ll, mc = len(lits), ml - 4
        blk.append((min(ll, 15) << 4) | min(mc, 15))
        if ll >= 15:
            blk += ext(ll - 15)
        blk += lits
        blk += bytes([off & 255, off >> 8])
        if mc >= 15:
            blk += ext(mc - 15)
        out += lits
        assert 1 <= off <= len(out)
        for _ in range(ml):
            out.append(out[-off])
    ll = len(tail_literals)
    blk.append(min(ll, 15) << 4)
    if ll >= 15:
        blk += ext(ll - 15)
    blk += tail_literals
    out += tail_literals
    return bytes(blk), bytes(out)


@pytest.mark.parametrize("misalign", [0, 1, 5, 8, 15])
def test_lz4_register_window_sequences(misalign, liblz4):
    """Hand-assembled high-ratio LZ4 blocks aimed at the direct decoder's register-window path: periods
    1/2/4/8 inside the sequence's literals (expanded from registers), the same periods reaching behind the
    literals, non-power-of-two and long periods (memory copy), literal counts 0..14 and 15+ (extension),
    match lengths on both sides of every internal threshold, length extensions that end inside / beyond the
    32-byte window.  Checked against liblz4 and the byte-serial expansion."""
    from gpu_util import gpu_decompress
    rng = np.random.default_rng(99)
    lens = [4, 5, 15, 18, 19, 20, 30, 31, 32, 33, 47, 63, 64, 65, 100, 273, 274, 500, 529, 4000, 7000, 9000]
    chunks, expect = [], []
    for period in [1, 2, 4, 8, 3, 5, 7, 12, 16, 40]:
        seqs = []
        first = True
        for i, ml in enumerate(lens):
            for ll in ([period, period + 1, 14] if period <= 13 else [14, 20]):
                lits = rng.integers(0, 256, ll, dtype=np.uint8).tobytes()
                if first and ll < period:
                    lits = rng.integers(0, 256, period, dtype=np.uint8).tobytes()
                seqs.append((lits, period, ml))
                first = False
            seqs.append((b"", period, ml))                       # no literals: period reaches behind
            seqs.append((rng.integers(0, 256, 2, dtype=np.uint8).tobytes(), period, ml))
        seqs.append((rng.integers(0, 256, 40, dtype=np.uint8).tobytes(), 8, 1000))   # 15+ literals
        blk, out = _lz4_block(seqs, b"tail-literals")
        assert len(out) >= 4 * len(blk)                          # routed to the direct decoder
        assert liblz4.decompress(blk, len(out)) == out
        chunks.append(blk); expect.append(out)
    outs, actual, status, _ = gpu_decompress(_codec("lz4"), chunks, [len(e) for e in expect], misalign=misalign)
    assert (status == 0).all(), status
    assert actual.tolist() == [len(e) for e in expect]
    for i, (o, e) in enumerate(zip(outs, expect)):
        assert o == e, i
