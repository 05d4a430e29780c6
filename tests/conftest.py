import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def emulib():
    """The kernel sources compiled for the SIMT emulator (tests/emu): test infrastructure
    that lets index logic and launch plans be checked without a GPU."""
    from causalstump_b200 import _lib
    here = os.path.join(ROOT, "tests", "emu")
    so = os.path.join(here, "libcs_emu.so")
    srcs = [os.path.join(here, f) for f in ("emu.cpp", "emu.h")]
    csrc = os.path.join(ROOT, "causalstump_b200", "csrc")
    srcs += [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["sh", os.path.join(here, "build_emu.sh")])
    return _lib.CsLib(so)


@pytest.fixture(scope="session")
def gpulib():
    """The product library on a real device; fails (not skips) if it cannot run."""
    from causalstump_b200 import _lib
    lib = _lib.load()
    lib.require_device()
    return lib
