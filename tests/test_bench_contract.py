"""bench.py's reference arm runs on CPU only: check the JSON contract of its line here (the GPU arm is exercised by
the driver on the B200 box)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    env = {**os.environ, "PHANT_BENCH_CPU_SECONDS": "1", "PHANT_BENCH_CPU_SAMPLE": "4096"}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "mpt_proofs_verified_per_sec" and line["unit"] == "proofs/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 2
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("synthetic account proofs")


def test_reference_arm_other_ranks_exit_quietly():
    env = {**os.environ, "RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
