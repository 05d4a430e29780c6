"""Memory-safety check without a GPU tool: the same kernel sources compiled for the SIMT emulator with
AddressSanitizer (compute-sanitizer is not available on the GPU pool).  Heap buffers stand in for device memory, so any
tile / panel / slab offset that leaves its buffer aborts the run."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_emulated_kernels_are_asan_clean(tmp_path):
    gxx = shutil.which("g++")
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not gxx or not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("g++ / libasan not available")
    lib = str(tmp_path / "libcs_emu_asan.so")
    emu = os.path.join(ROOT, "tests", "emu")
    csrc = os.path.join(ROOT, "causalstump_b200", "csrc")
    subprocess.check_call([gxx, "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-fsanitize=address", "-fno-omit-frame-pointer",
                           "-DCS_EMU", "-x", "c++", "-I" + emu, "-I" + csrc, os.path.join(csrc, "api.cu"), "-x", "c++",
                           os.path.join(emu, "emu.cpp"), "-o", lib])
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1")
    r = subprocess.run([sys.executable, os.path.join(emu, "asan_workload.py"), lib], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0 and "asan workload done" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
    assert "AddressSanitizer" not in r.stderr
