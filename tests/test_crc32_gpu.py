"""-m gpu: standard CRC-32 low-level API (nvcompBatchedCRC32Async, include/nvcomp/crc32.h; reference
CHANGELOG.md:51, examples/standard_crc_checksum.cpp:94-104 checks it against boost::crc_32_type == zlib.crc32)."""
import ctypes as C
import zlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _crc_batch(chunks, misalign=0):
    import nvcomp_b200
    from nvcomp_b200.batched import make_batch
    lib = nvcomp_b200.load()
    fn = lib.nvcompBatchedCRC32Async
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    b = make_batch(chunks, misalign=misalign)
    out = torch.zeros(len(chunks), dtype=torch.int32, device="cuda")
    st = fn(b.ptrs.data_ptr(), b.sizes.data_ptr(), len(chunks), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert st == 0
    torch.cuda.synchronize()
    return out.cpu().numpy().view(np.uint32).tolist()


@pytest.mark.parametrize("misalign", [0, 1, 3, 7])
def test_crc32_matches_zlib(misalign):
    rng = np.random.default_rng(12)      # the reference example: random bytes, random chunk sizes in [1, 1024)
    chunks = [rng.integers(0, 256, int(n), dtype=np.uint8).tobytes() for n in rng.integers(1, 1024, 300)]
    chunks += [b"", b"a", b"123456789", bytes(65536), rng.integers(0, 256, 65536, dtype=np.uint8).tobytes(),
               rng.integers(0, 256, 1 << 20, dtype=np.uint8).tobytes(), rng.integers(0, 256, 99999, dtype=np.uint8).tobytes()]
    got = _crc_batch(chunks, misalign)
    want = [zlib.crc32(c) & 0xffffffff for c in chunks]
    assert got == want
    assert zlib.crc32(b"123456789") == 0xCBF43926       # the CRC-32 check value
