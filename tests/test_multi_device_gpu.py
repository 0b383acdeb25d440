"""-m gpu (needs >= 2 GPUs, skipped otherwise): ONE process drives two devices -- the reference's benchmark_allgather
pattern (benchmarks/benchmark_allgather.cpp:359-368).  Function attributes (opt-in dynamic shared memory, carveout),
the CRC constant tables and the LZ decoders' side stream are per device; round 1 guarded them with one per-process
flag, so the first launch on a second GPU failed."""
import zlib

import numpy as np
import pytest
import torch

from conftest import sample_inputs

pytestmark = pytest.mark.gpu
INPUTS = sample_inputs()


def _roundtrip_all_codecs(dev):
    from gpu_util import gpu_compress, gpu_decompress
    from nvcomp_b200._lib import BitcompOpts, CascadedOpts
    from nvcomp_b200.batched import Codec
    raws = [INPUTS[n] for n in ("sorted_i64", "price_walk", "runlength_i32", "lowentropy", "random_64k")]
    with torch.cuda.device(dev):
        for codec in (Codec("Cascaded", opts=CascadedOpts(4096, 6, 1, 1, 1)), Codec("Bitcomp", opts=BitcompOpts(0, 7)),
                      Codec("ANS"), Codec("LZ4"), Codec("Snappy")):
            comps, _ = gpu_compress(codec, raws)
            outs, actual, status, _ = gpu_decompress(codec, comps, [len(r) for r in raws])
            assert (status == 0).all() and outs == raws, (dev, codec.fmt)


def test_every_codec_on_both_devices_from_one_process():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    _roundtrip_all_codecs(0)
    _roundtrip_all_codecs(1)      # 96 KB-smem Cascaded kernels, LZ side stream, all on a device that was never set up
    _roundtrip_all_codecs(0)


def test_crc32_on_second_device():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    from test_crc32_gpu import _crc_batch
    chunks = [INPUTS["text"], INPUTS["random_64k"], b"123456789"]
    with torch.cuda.device(1):
        assert _crc_batch(chunks) == [zlib.crc32(c) & 0xffffffff for c in chunks]
