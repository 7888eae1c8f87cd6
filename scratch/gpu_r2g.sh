#!/bin/bash
TAG=${1:-r2g}
O=gpurun_out
mkdir -p $O
B="python bench.py --no-cpu-baseline --steps 100 --warmup 5 --no-ddp-leg --no-reference-gpu"
$B > $O/${TAG}_base.json 2>/dev/null
PD_STREAM_FWD_MINB=5 $B > $O/${TAG}_fwd5.json 2>/dev/null
PD_STREAM_FWD_MINB=6 $B > $O/${TAG}_fwd6.json 2>/dev/null
PD_STREAM_FWD_MINB=5 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k "cfg2" > $O/${TAG}_pytest_fwd5.log 2>&1; echo "fwd5 tests rc=$?"
PD_STREAM_FWD_MINB=6 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k "cfg2" > $O/${TAG}_pytest_fwd6.log 2>&1; echo "fwd6 tests rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2g_*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], "%.4f ms  %.0f img/s"%(d["ms_per_step"], d["value"]), {k:round(v,4) for k,v in d["roofline"]["all_kernels_ms"].items()}, "e2e", round(d["e2e"]["value"]), round(d["e2e"]["other_transport"]["value"]))
    except Exception as e: print(f, "ERR", e)
PY
