"""The C-ABI library builds, loads and exports every symbol include/cs_b200.h declares.
No compute is possible without a GPU: device-less calls must fail loudly."""
import os
import re

import pytest

from causalstump_b200 import _lib
from causalstump_b200.csrc import build as csbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def product_lib():
    csbuild.build()
    return _lib.CsLib(_lib.DEFAULT_LIB)


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cs_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(product_lib):
    decl = declared_symbols()
    assert len(decl) >= 28
    for name in decl:
        assert hasattr(product_lib.dll, name), name
    assert sorted(_lib.EXPORTED_SYMBOLS) == decl  # the ctypes table covers the whole header


def test_version_and_error_channel(product_lib):
    assert product_lib.dll.cs_version() == 100
    assert isinstance(product_lib.dll.cs_last_error(), bytes)


def test_no_cpu_fallback(product_lib):
    import numpy as np
    if product_lib.dll.cs_device_count() > 0:
        pytest.skip("a CUDA device is present; the device-less failure path is not reachable")
    with pytest.raises(_lib.CsError) as ei:
        product_lib.kernmat(np.zeros((4, 1)), np.zeros((4, 1)), np.zeros(4), np.zeros(4), [0.0], [0.0], 0.0, 0.0)
    assert ei.value.code == -4 and "no CPU fallback" in str(ei.value)
    with pytest.raises(_lib.CsError):
        _lib.Fit(product_lib, 0, np.zeros(4), np.zeros((4, 1)), np.zeros(4), None, np.zeros(6))
    with pytest.raises(_lib.CsError):
        product_lib.require_device()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "causalstump_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, fn
                assert "tests/emu" not in txt.replace("tests/emu only", "").replace("(tests/emu", "(") or fn in ("cs_rt.h", "api.cu", "cs_dev.cuh"), fn


def test_sass_uses_the_fp64_tensor_path():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    csbuild.build()
    sass = subprocess.run([cuobjdump, "-sass", _lib.DEFAULT_LIB], capture_output=True, text=True).stdout
    assert "DMMA.8x8x4" in sass and "LDGSTS" in sass
    assert "sm_100a" in subprocess.run([cuobjdump, "-lelf", _lib.DEFAULT_LIB], capture_output=True, text=True).stdout
