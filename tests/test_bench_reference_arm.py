"""The reference arm of bench.py needs no GPU: run it on a tiny sample and check the contract of its JSON line."""
import json
import os
import subprocess
import sys

ROOT_ = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT_ not in sys.path:
    sys.path.insert(0, ROOT_)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "ex", "--steps", "2",
           "--warmup", "1", "--cpu-sample-windows", "600", "--cpu-walk-seconds", "1.0"]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "cbow_context_windows_per_sec"
    assert line["unit"] == "windows/s" and line["higher_is_better"] is True and line["value"] > 0
    # the UNMODIFIED reference script when it is staged (oracle/_ref, or /root/reference here), else the oracle port
    from oracle import ref_import
    kind = "reference" if ref_import.available() else "port"
    assert line["cpu_baseline"]["kind"] == kind and line["cpu_baseline"]["value"] == line["value"]
    assert line["walk"]["kind"] == kind and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["walk"]["value"] > 0 and line["walk"]["unit"] == "steps/s"
    # ranks other than 0 of a torchrun launch exit quietly
    r2 = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=60, env=dict(env, RANK="1", WORLD_SIZE="2"))
    assert r2.returncode == 0 and r2.stdout.strip() == ""
