"""bench.py contract pieces that can be checked without a GPU: the reference arm's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "3", "--warmup", "3"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "chunks/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("chunks/sec") and line["value"] > 0 and line["steps"] == 3
    assert line["e2e"] == {"value": line["value"], "unit": "chunks/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert "workload" in line["config"] and line["gpu_launches"] == 0


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--gpus", "2", "--steps", "3"], capture_output=True, text=True, timeout=120,
                         env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
