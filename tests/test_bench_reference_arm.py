"""The reference arm of bench.py needs no GPU: run it on a tiny sample and check the contract of its JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "ex", "--steps", "2",
           "--warmup", "1", "--cpu-sample-windows", "600", "--cpu-sample-starts", "40"]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "cbow_context_windows_per_sec"
    assert line["unit"] == "windows/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["walk"]["value"] > 0 and line["walk"]["unit"] == "steps/s"
    # ranks other than 0 of a torchrun launch exit quietly
    r2 = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=60, env=dict(env, RANK="1", WORLD_SIZE="2"))
    assert r2.returncode == 0 and r2.stdout.strip() == ""
