"""Helpers shared by the -m gpu parity tests (all calls go through the C ABI via nvcomp_b200.batched)."""
import numpy as np
import torch

from nvcomp_b200.batched import Batch, Codec, empty_batch, make_batch


def gpu_decompress(codec: Codec, comp_chunks, caps, misalign=0, want_actual=True, want_status=True):
    """comp_chunks: list of bytes; caps: list of output capacities.  Returns (outputs, actual, status)."""
    comp = make_batch(comp_chunks, misalign=misalign)
    out = empty_batch(len(comp_chunks), max(max(caps), 1) if len(caps) else 1, misalign=misalign, fill=0xA5)
    out.sizes = torch.tensor(list(caps), dtype=torch.int64, device="cuda")
    actual, status = codec.decompress(comp, out, want_actual=want_actual, want_status=want_status)
    torch.cuda.synchronize()
    a = actual.cpu().numpy() if actual is not None else None
    s = status.cpu().numpy() if status is not None else None
    sizes = a if a is not None else np.asarray(caps)
    return out.to_host(sizes), a, s, out


def gpu_compress(codec: Codec, raw_chunks, misalign=0):
    inp = make_batch(raw_chunks, misalign=misalign)
    comp = codec.compress(inp, max_chunk=max([len(c) for c in raw_chunks] + [1]))
    torch.cuda.synchronize()
    sizes = comp.sizes.cpu().numpy()
    return comp.to_host(sizes), comp
