"""GPU (-m gpu): bench.py exactly as the driver launches it at N = 1 (every leg on: value, e2e in both transports, roofline,
ddp leg, reference_gpu leg, cpu_baseline) prints ONE JSON line with the contract's keys.  Guards the ordering of the legs: the
CPU arm patches `.cuda()` while it runs the reference and must not leak that into the GPU legs."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def test_default_bench_line_has_every_leg():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "3"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:2000]
    j = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in j, k
    assert j["n_gpus"] == 1 and j["steps"] == 20 and j["value"] > 0 and j["gpu_launches"] > 0
    assert j["e2e"]["value"] > 0 and j["e2e"]["h2d_bytes_per_step"] > 0 and j["e2e"]["d2h_bytes_per_step"] == 4
    rf = j["roofline"]
    assert rf["bound"] == "hbm" and 0 < rf["frac"] < 1.2 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    cb = j["cpu_baseline"]
    assert cb and cb["value"] > 0 and cb["cores"] >= 1 and cb["kind"] in ("reference", "port")
    assert j["ddp"]["value"] > 0 and j["ddp"]["n_gpus"] == 1
    assert "reference_gpu" in j and ("value" in j["reference_gpu"] or "unavailable" in j["reference_gpu"])
    if "value" in j["reference_gpu"]:
        # the reference's loss on the same batch agrees with ours (a sanity check of the comparison, not a parity test)
        assert abs(j["reference_gpu"]["loss"] - j["loss"]) < 1e-4
