#!/bin/bash
TAG=${1:-r2i}
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/${TAG}_pytest.log
B="python bench.py --no-cpu-baseline --steps 200 --warmup 10 --no-ddp-leg --no-reference-gpu"
for i in 1 2; do
  $B > $O/${TAG}_hint_fused_$i.json 2>/dev/null
  $B --no-fuse-bwd > $O/${TAG}_hint_unfused_$i.json 2>/dev/null
  PD_STREAM_NO_L2_HINT=1 $B > $O/${TAG}_nohint_fused_$i.json 2>/dev/null
done
for c in cfg3 cfg5; do
  python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline --no-ddp-leg --no-reference-gpu > $O/${TAG}_hint_$c.json 2>/dev/null
  PD_STREAM_NO_L2_HINT=1 python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline --no-ddp-leg --no-reference-gpu > $O/${TAG}_nohint_$c.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2i_*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], "%.4f ms  %.0f img/s"%(d["ms_per_step"], d["value"]), {k:round(v,4) for k,v in d["roofline"]["all_kernels_ms"].items()}, "frac", round(d["roofline"]["frac"],3))
    except Exception as e: print(f, "ERR", e)
PY
