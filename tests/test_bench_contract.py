"""bench.py contract that can be checked without a GPU: the reference arm prints exactly ONE line on stdout, a JSON
object with the agreed keys (metric / unit / config of BASELINE.json, cpu_baseline, e2e with zero copy bytes)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_json_line(oracle_port):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--rows", "24", "--cols", "20", "--cpu-seconds", "0.5"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pixel inversions/s" and d["unit"] == "px/s"
    assert d["higher_is_better"] is True and d["scaling"] == "strong" and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--rows", "24", "--cols", "20"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
