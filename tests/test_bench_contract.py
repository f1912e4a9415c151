"""The driver-facing contract of `bench.py --impl reference` (the CPU arm), checked without a GPU on a small grid:
one JSON line with the metric / unit / config of the GPU arm, `impl`, `cpu_baseline` {value, unit, cores, kind, sample}
and `e2e` {value, unit, h2d_bytes_per_step, d2h_bytes_per_step}; under torchrun only rank 0 works."""
import json
import os
import subprocess
import sys

from tests.conftest import ROOT


def _run(extra_env=None, *args):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--nx", "32",
                           "--cpu-nx", "32", "--steps", "2", "--warmup", "1", *args],
                          capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "apply_inverse_per_s" and d["unit"] == "1/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["ms_per_step"] - 1e3 / d["value"]) < 1e-9 * d["ms_per_step"]
    assert "workload" in d["config"] and "32^3" in d["config"]["workload"]
    assert d["config"]["same_config"] is True and d["config"]["extrapolated"] is False
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["unit"] == "1/s" and cb["value"] == d["value"]
    assert "F-matrix ordering" in cb["sample"] and "32^3" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "1/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    g = d["gmres"]
    assert g["converged"] and g["iterations"] > 0 and g["explicit_rel_residual"] < 1e-6 and len(g["history_first15"]) > 1


def test_reference_arm_other_ranks_exit_without_work():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2", "--no-solve")
    assert r.returncode == 0 and r.stdout.strip() == ""
