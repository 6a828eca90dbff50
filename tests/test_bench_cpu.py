"""CPU: the reference arm of bench.py (`--impl reference`) keeps its contract -- one JSON line, the step / warm-up counts it
actually ran, `cpu_baseline.kind` saying whether the unmodified reference modules (baseline/_ref or /root/reference) or the
oracle port were timed, zero transfer bytes.  A tiny batch keeps it to a few seconds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_a_truthful_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--batch", "4"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "samples/s" and line["higher_is_better"] is True
    assert line["steps"] == 2 and line["warmup"] == 1                      # what was run, not merely what was asked for
    assert line["value"] > 0 and abs(line["value"] - 4 / (line["ms_per_step"] / 1e3)) < 0.05 * line["value"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and "B=4" in cb["sample"]
    from oracle import ref_import
    assert cb["kind"] == ("reference" if ref_import.available() else "port")
    assert line["e2e"] == {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["global_batch"] == 4 and line["dtype"] == "fp32"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
