"""CPU: the driver-facing contract of bench.py that can be checked without a GPU — the reference arm prints exactly one
JSON line with the contract's keys (same metric / unit / config as our arm), and our arm refuses to run without CUDA
(there is no CPU fallback to time by accident)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--ref-evals", "2", "--queries", "16384"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("generated frames/sec") and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["scaling"] == "strong" and d["vs_baseline"] is None
    # ms_per_step is the MEASURED duration of a (bounded-sample) step: steps x ms_per_step is time really spent
    assert d["value"] > 0 and d["steps"] == 1 and abs(d["ms_per_step"] * d["steps"] - 1000.0 * d["timed_s"]) < 1.0
    assert d["timed_s"] < d["wall_s"] and d["extrapolated"] is True and 0 < d["sample_fraction_of_frame"] < 1
    assert abs(d["ms_per_global_step_extrapolated"] * d["value"] - 64 * 1000.0) < 1e-3 * 64 * 1000.0
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "oracle port" in cb["sample"]
    assert d["config"]["workload"].startswith("configs[2]") and d["config"]["global_frames"] == 64
    assert d["config"]["frames_per_gpu"] == 64 and d["config"]["queries_per_frame"] == 16384
    assert d["gpu_launches"] == 0


def test_our_arm_needs_cuda():
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
