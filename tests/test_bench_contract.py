"""bench.py's reference arm (`--impl reference`) runs without a GPU: one bounded sample of the CPU forward (oracle port,
the checker -- the one place bench.py may execute oracle/), printed as ONE JSON line with the contract's keys; under a
multi-rank launch only rank 0 works and prints."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", env_extra.get("WORLD_SIZE", "1"),
                           "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_one_contract_line_on_cpu():
    res = _run({})
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "episodes/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("episodes/sec 5-way 5-shot Meta-FCOS R-50")
    assert d["steps"] == 1 and d["warmup"] == 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("5-way 5-shot Meta-FCOS R-50 FPN, 8 query images") and "model" not in d["config"]
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1000.0) < 1e-6
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and "800x1333" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_without_work():
    res = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0 and res.stdout.strip() == ""
