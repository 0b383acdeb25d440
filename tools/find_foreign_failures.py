#!/usr/bin/env python3
"""Developer tool: compress a dataset on the host with liblz4 (default / HC-12) or pyarrow-snappy, decode on the GPU,
and dump every chunk whose status / size / bytes are wrong to gpurun_out/fail/ for the host warp emulator."""
import ctypes as C
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    kind, dataset, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
    data = bench.gen_data(dataset, n)
    if len(sys.argv) > 4:                       # first run the own-compressor workload, like bench.py does
        w0 = bench.Workload(kind, dataset, n, data=data)
        bench.time_decode(w0, 5, 3)
        del w0
    lz4 = C.CDLL("liblz4.so.1")
    lz4.LZ4_compress_default.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    lz4.LZ4_compress_HC.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
    cap = 65536 + 65536 // 255 + 64

    def mk(hc):
        def f(raw):
            buf = C.create_string_buffer(cap)
            n_ = lz4.LZ4_compress_HC(raw, buf, len(raw), cap, hc) if hc else lz4.LZ4_compress_default(raw, buf, len(raw), cap)
            return buf.raw[:n_]
        return f
    producers = {"lz4": [("default", mk(0)), ("hc12", mk(12))]}
    if kind == "snappy":
        import pyarrow as pa
        codec = pa.Codec("snappy")
        producers["snappy"] = [("pyarrow", lambda raw: codec.compress(raw).to_pybytes())]
    os.makedirs("gpurun_out/fail", exist_ok=True)
    for name, fn in producers[kind]:
        with ThreadPoolExecutor(max_workers=64) as ex:
            chunks = list(ex.map(lambda i: fn(data[i].tobytes()), range(n)))
        w = bench.Workload(kind, dataset, n, data=data, comp_chunks=chunks)
        for rep in range(4):
            w.reset_outputs()
            for _ in range(1 if rep == 0 else 4):     # rep > 0: back-to-back launches, as the bench warm-up does
                w.launch(torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            st = w.status.cpu().numpy(); ac = w.actual.cpu().numpy()
            out = w.out.slab[: w.total].cpu().numpy().reshape(n, 65536)
            bad = [i for i in range(n) if st[i] != 0 or ac[i] != 65536 or not np.array_equal(out[i], data[i])]
            print(name, "rep", rep, "bad chunks:", bad[:20], "count", len(bad), flush=True)
            for i in bad[:4]:
                open(f"gpurun_out/fail/{kind}_{name}_{i}.comp", "wb").write(chunks[i])
                open(f"gpurun_out/fail/{kind}_{name}_{i}.raw", "wb").write(data[i].tobytes())
                open(f"gpurun_out/fail/{kind}_{name}_{i}.gpu", "wb").write(out[i].tobytes())
                print("   chunk", i, "status", st[i], "actual", ac[i], "clen", len(chunks[i]),
                      "first diff", int(np.argmax(out[i] != data[i])) if ac[i] == 65536 else -1)


if __name__ == "__main__":
    main()
