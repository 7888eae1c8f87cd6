#!/bin/bash
# generic A/B over a tuning environment variable: gpu_ab_knob.sh TAG VAR "v0 v1 ..." "cfgs"
TAG=$1; VAR=$2; VALS=$3; CFGS=${4:-cfg2}
O=gpurun_out
mkdir -p $O
B="--steps 40 --warmup 5 --no-cpu-baseline --no-ddp-leg --no-reference-gpu"
for cfg in $CFGS; do
for rep in 1 2; do
for v in $VALS; do
  env $VAR=$v python bench.py --config $cfg $B > $O/${TAG}_${cfg}_$v.json 2>$O/${TAG}_err.log || tail -n 3 $O/${TAG}_err.log
  python - <<PY
import json
try:
    d=json.load(open("$O/${TAG}_${cfg}_$v.json")); print("$cfg $VAR=$v", "%.4f ms"%d["ms_per_step"], {k:round(x,4) for k,x in d["roofline"]["all_kernels_ms"].items()})
except Exception as e: print("failed", e)
PY
done
done
done
