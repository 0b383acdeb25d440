"""-m gpu: high-level interface (nvcomp::*Manager, create_manager, checksum policies) -- runs the C++ test
binary built from tests/cpp/hlif_test.cu (HLIF is a C++ API; the reference's callers are C++)."""
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_hlif_cpp():
    exe = os.path.join(ROOT, "build", "tests", "hlif_test")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", ROOT, "build/tests/hlif_test"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])
    assert "hlif_test ok" in r.stdout
