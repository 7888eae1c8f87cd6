#!/bin/bash
TAG=${1:-r2o}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bf16" > $O/${TAG}_pytest_bf16.log 2>&1; echo "bf16 tests rc=$?"; tail -n 25 $O/${TAG}_pytest_bf16.log | cut -c1-220
timeout 900 python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/${TAG}_pytest.log
B="python bench.py --no-cpu-baseline --steps 100 --warmup 5 --no-ddp-leg --no-reference-gpu"
timeout 300 $B > $O/${TAG}_cfg2_fp32.json 2>/dev/null
timeout 300 $B --storage bf16 > $O/${TAG}_cfg2_bf16.json 2> $O/${TAG}_cfg2_bf16.err
timeout 300 $B --config cfg3 --steps 30 > $O/${TAG}_cfg3_fp32.json 2>/dev/null
timeout 300 $B --config cfg3 --steps 30 --storage bf16 > $O/${TAG}_cfg3_bf16.json 2> $O/${TAG}_cfg3_bf16.err
tail -n 3 $O/${TAG}_cfg2_bf16.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2o_*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], "%.4f ms  %.0f img/s"%(d["ms_per_step"], d["value"]), {k:round(v,4) for k,v in d["roofline"]["all_kernels_ms"].items()}, "frac", {k:round(v,3) for k,v in d["roofline"]["all_kernels_frac"].items()}, "loss", d["loss"])
    except Exception as e: print(f, "ERR", e)
PY
