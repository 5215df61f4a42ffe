"""bench.py contract on CPU: the reference arm (`--impl reference`: the unmodified reference staged in oracle/_ref or, when
absent, the oracle port, on host cores) prints ONE JSON line with the keys the driver reads.  (The B200 arm needs a GPU; its
line is checked by the driver's own run.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "train", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "captions_per_sec" and d["unit"] == "captions/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("train_step")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and abs(cb["value"] - d["value"]) < 1e-6 * d["value"]
    assert "bounded sample of 16 captions" in cb["sample"] and "bounded sample of 16 captions" in d["config"]["workload"]
    assert d["config"]["device"] == "cpu" and d["dtype"] == "f32"
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0 and e["unit"] == d["unit"]
    assert abs(e["value"] - d["value"]) < 1e-6 * d["value"]
