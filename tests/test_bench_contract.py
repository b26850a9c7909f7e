"""bench.py's reference arm (`--impl reference`: the reference's own classes on the host cores when the reference tree or its
staged copy is on the machine, the oracle port otherwise) runs without a GPU: check that it prints exactly one JSON line on
stdout with the keys of the driver's contract, for both workloads."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


@pytest.mark.parametrize("model,nodes", [("mus", 3000), ("remus", 800)])
def test_reference_arm_prints_one_json_line(model, nodes):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1",
           "--model", model, "--nodes", str(nodes), "--cpu-sample-nodes", str(nodes), "--weights", "init"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["metric"] == "rollout_steps_per_s" and d["unit"] == "steps/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    from conftest import HAVE_REFERENCE
    assert d["cpu_baseline"]["kind"] == ("reference" if HAVE_REFERENCE else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["gpu_eager"] is None                       # no GPU on this machine
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_reference_arm_fit_and_shipped_weights():
    """Three timed steps switch the straight-line fit on; hidden 128 / 3 scales picks the staged shipped checkpoint."""
    from conftest import HAVE_REFERENCE
    if not HAVE_REFERENCE:
        pytest.skip("no reference tree")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "0",
           "--nodes", "20000", "--cpu-sample-nodes", "2000"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    d = json.loads(res.stdout.strip().splitlines()[-1])
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and len(cb["seconds_per_step_by_nodes"]) == 3 and cb["sample_fraction"] == 0.1
    assert "shipped 3S-GNN" in d["config"]["weights"] or "no staged checkpoint" in d["config"]["weights"]
    assert abs(d["ms_per_step"] - 1e3 * cb["seconds_per_step_by_nodes"]["2000"]) < 1e-6      # what was really timed
