"""bench.py's reference arm (`--impl reference`) runs on the host cores alone, so its half of the bench contract can be
checked without a GPU: one JSON line with the keys the driver reads, the oracle as the thing timed (`cpu_baseline.kind`
"port"), no device copies in `e2e`; under a multi-rank launch only rank 0 works and prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c1", "--steps", "1",
                           "--warmup", "1"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_the_contract_line():
    p = _run()
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                   # ONE JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"].startswith("Mpaths/s") and d["unit"] == "Mpaths/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["scaling"] in ("weak", "strong") and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic"
    assert "64^3" in d["config"]["workload"] and d["config"]["name"] == "c1" and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_exit_without_work():
    p = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29533"})
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stdout.strip() == ""
