"""bench.py's own arm on a small configuration: one JSON line carrying every key of the measurement
contract (value / e2e / roofline / cpu_baseline / gpu_launches / clocks / config.workload)."""
from __future__ import annotations

import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_native_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--trees", "9472", "--iterations", "300",
                          "--steps", "2", "--warmup", "1", "--cpu-repeats", "1", "--leaf-batch", "2"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[-2000:]
    line = json.loads(lines[0])
    assert "impl" not in line or line["impl"] != "reference"
    assert line["metric"] == "crossingstate_env_steps_per_s" and line["unit"] == "env-steps/s"
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1
    assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["dtype"] == "f64" and line["data"] == "synthetic"
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["iterations_per_s"] > 0
    assert "workload" in line["config"] and line["config"]["trees_per_gpu"] == 9472
    assert line["gpu_launches"] >= 2 * line["steps"]          # reset + search per step, counted by the engine
    e2e = line["e2e"]
    assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0
    assert e2e["value"] != line["value"]                       # measured on its own, not copied
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert line["leaf_batched"]["rollouts_per_leaf"] == 2 and line["leaf_batched"]["env_steps_per_s"] > 0
    assert line["rollout_c3"]["env_steps_per_s"] > 0
