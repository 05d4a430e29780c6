"""bench.py contract checks that need no GPU: the reference arm's JSON line (keys the driver reads) and that ranks other
than 0 stay silent.  The CPU timing itself is stubbed (the real one takes about a minute per step)."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_reference_arm_line(monkeypatch, capsys):
    monkeypatch.setattr(bench, "cpu_eval_seconds", lambda inp, budget: (100.0, "stub sample"))
    monkeypatch.setattr(bench, "make_inputs", lambda n, p, r, kind="GP": {})
    args = types.SimpleNamespace(steps=2, warmup=1, gpus=4)
    assert bench.run_reference(args, rank=1, world=4) == 0 and capsys.readouterr().out == ""   # only rank 0 works and prints
    assert bench.run_reference(args, rank=0, world=4) == 0
    line = json.loads(capsys.readouterr().out.strip())
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "impl", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == "evals/s" and line["n_gpus"] == 4
    assert abs(line["value"] - 0.01) < 1e-12 and line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "workload" in line["config"]
    assert line["vs_baseline"] is None and line["dtype"] == "f64" and line["higher_is_better"] is True


def test_ours_fails_loudly_without_a_gpu():
    """the product arm never falls back to the CPU: without a CUDA device it raises"""
    import pytest
    from causalstump_b200 import _lib
    lib = _lib.load()
    if lib.dll.cs_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.CsError):
        lib.require_device()
