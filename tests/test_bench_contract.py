"""bench.py's reference arm (`--impl reference`) runs without a GPU: check the JSON contract of the line it prints
(one line on stdout; metric / unit / config shared with our arm; `impl`, `cpu_baseline` and `e2e` as the harness expects)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", *extra],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, **(env or {})))
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    stdout = _run("--cpu-size", "24")
    lines = [ln for ln in stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "Mcell-updates/s" and d["unit"] == "Mcell-updates/s" and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "24^3" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun (N > 1) rank 0 alone runs the CPU restatement; the other ranks exit 0 without work or output."""
    stdout = _run("--cpu-size", "24", "--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert stdout.strip() == ""


def test_reference_arm_compressible_workload():
    d = json.loads(_run("--workload", "supercell").strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["value"] > 0 and "compressible" in d["config"]["workload"]
