"""CPU: the C-ABI shared library loads and exports every symbol include/nvcomp/*.h declares
(no compute calls -- there is no GPU here); host-only size queries behave."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, _ensure_built


def _declared_symbols():
    syms = []
    inc = os.path.join(ROOT, "include", "nvcomp")
    for fn in sorted(os.listdir(inc)):
        if not fn.endswith(".h"):
            continue
        txt = open(os.path.join(inc, fn)).read()
        syms += re.findall(r"nvcompStatus_t\s+(nvcomp\w+)\s*\(", txt)
    return sorted(set(syms))


def test_library_exports_every_declared_symbol():
    _ensure_built()
    from nvcomp_b200 import lib_path
    lib = C.CDLL(lib_path())
    syms = _declared_symbols()
    assert len(syms) >= 16
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_host_only_size_queries():
    from nvcomp_b200.batched import Codec
    lz4 = Codec("LZ4")
    assert lz4.compress_get_max_output_chunk_size(65536) >= 65536 + 65536 // 255 + 16
    assert lz4.decompress_get_temp_size(10000, 65536) >= 0
    assert lz4.compress_get_temp_size(10000, 65536) >= 0
    sn = Codec("Snappy")
    assert sn.compress_get_max_output_chunk_size(65536) >= 32 + 65536 + 65536 // 6
    from nvcomp_b200.batched import NvcompError
    with pytest.raises(NvcompError):
        lz4.compress_get_max_output_chunk_size((1 << 24) + 1)


def test_no_cpu_fallback_in_product():
    """The product package must not import or reference the oracle (oracle/ is test infrastructure)."""
    pkg = os.path.join(ROOT, "nvcomp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in txt and "oracle_" not in txt, os.path.join(dirpath, f)


def test_call_logging_env(tmp_path):
    """NVCOMP_LOG_LEVEL=3 logs every low-level call to NVCOMP_LOG_FILE (reference README.md:79-88)."""
    import subprocess
    import sys
    log = tmp_path / "nv.log"
    code = ("from nvcomp_b200.batched import Codec\n"
            "c = Codec('Snappy')\n"
            "c.decompress_async(None, None, None, None, 0, None, 0, None, None, None)\n")
    env = dict(os.environ, NVCOMP_LOG_LEVEL="3", NVCOMP_LOG_FILE=str(log), PYTHONPATH=ROOT)
    subprocess.run([sys.executable, "-c", code], check=True, env=env, cwd=ROOT)
    assert "nvcompBatchedSnappyDecompressAsync(batch_size=0" in log.read_text()
