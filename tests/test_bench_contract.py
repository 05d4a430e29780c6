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
    monkeypatch.setattr(bench, "ref_available", lambda: False)   # compiled reference absent -> oracle port
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


def test_reference_arm_uses_the_compiled_reference(monkeypatch, capsys):
    """with oracle/_ref present the arm times the reference's own C++ (kind "reference") and reports the port beside it"""
    monkeypatch.setattr(bench, "cpu_eval_seconds", lambda inp, budget: (100.0, "stub port"))
    monkeypatch.setattr(bench, "ref_eval_seconds", lambda inp, budget: (200.0, "stub model"))
    monkeypatch.setattr(bench, "ref_available", lambda: True)
    full = dict(kernmat=60.0, invkernel=10.0, grad=179.0, mu_solution=1.0, log_det_alone=4.0, invkernel_2threads=30.0,
                log_det_2threads=14.0, lapack="stub lapack", lml=-1.0)
    monkeypatch.setattr(bench, "ref_full_eval", lambda inp, threads: dict(full))
    monkeypatch.setattr(bench, "make_inputs", lambda n, p, r, kind="GP": {})
    assert bench.run_reference(types.SimpleNamespace(steps=20, warmup=5, gpus=1), rank=0, world=1) == 0
    line = json.loads(capsys.readouterr().out.strip())
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and abs(line["value"] - 0.004) < 1e-12 and cb["value"] == line["value"]
    assert abs(cb["oracle_port"]["value"] - 0.01) < 1e-12 and line["e2e"]["value"] == line["value"]
    # one honest full-size evaluation: the line reports the steps it really ran, so ms_per_step x steps is the run time
    assert line["steps"] == 1 and line["warmup"] == 0 and line["steps_requested"] == 20 and abs(line["ms_per_step"] - 250e3) < 1e-6
    assert abs(cb["omp_num_threads_2"]["seconds"] - 280.0) < 1e-9 and abs(cb["model"]["ratio_to_measured"] - 0.8) < 1e-12
    assert line["config"] == bench.workload_config()   # same `config` object as the product arm at N = 1


def test_no_roofline_fraction_above_the_bracket():
    import pytest
    bench.check_fracs({"roofline": {"frac": 0.84}, "x": {"y": {"frac": 1.1}}})
    with pytest.raises(AssertionError):
        bench.check_fracs({"roofline_elementwise": {"gram": {"frac": 2.7}}})


def test_reference_timing_sample_on_a_small_case():
    """ref_eval_seconds runs the compiled reference end to end (LAPACK hooked in) and extrapolates a positive time"""
    import pytest
    from oracle import ref_binding as rb
    if not rb.available():
        pytest.skip("oracle/_ref not built")
    inp = bench.make_inputs(512, 25, 0)
    secs, desc = bench.ref_eval_seconds(inp, 1.0)   # n_s = 128
    assert secs > 0 and "n_s=128" in desc and "openblas" in desc.lower()


def test_ours_fails_loudly_without_a_gpu():
    """the product arm never falls back to the CPU: without a CUDA device it raises"""
    import pytest
    from causalstump_b200 import _lib
    lib = _lib.load()
    if lib.dll.cs_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.CsError):
        lib.require_device()
