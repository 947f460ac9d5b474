"""bench.py's reference arm (the reference's own C path, oracle/_ref, timed alone) prints one
well-formed JSON line.  Runs the cfg3 arm with a tiny bounded problem so that it fits the CPU suite."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg3",
                          "--steps", "1", "--warmup", "0", "--ref-sample-n", "256"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "points/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["cpu_baseline"]["kind"] == "reference" and j["cpu_baseline"]["cores"] == 1
    assert "EXTRAPOLATED" in j["cpu_baseline"]["sample"] and j["cpu_baseline"]["run_n"] == 256
    assert j["e2e"] == {"value": j["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["config"]["n"] == 1000000 and j["n_gpus"] == 1 and j["vs_baseline"] is None
