"""bench.py's reference arm prints the driver's JSON contract (CPU only: the arm times the oracle, no GPU involved)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, env=None):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--quick", "--steps", "1", "--warmup", "1",
           "--ref-batch", "500"] + extra
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=e, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in p.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_line():
    lines = _run([])
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"):
        assert key in d, key
    assert d["impl"] == "reference" and d["vs_baseline"] is None and d["higher_is_better"] is True
    assert d["unit"] == "read pairs/s" and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb


def test_reference_arm_other_ranks_stay_silent():
    # under torchrun only rank 0 runs the CPU arm; the other ranks exit 0 without work
    assert _run(["--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
