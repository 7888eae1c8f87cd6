#!/bin/bash
# ring-shape sweep of the streamed kernels (planes per stage x stages) through the load-time tuning variables
# usage: gpu_ring_sweep.sh TAG "hs nst" "hs nst" ...
TAG=${1:-r2v}; shift
O=gpurun_out
mkdir -p $O
B="--steps 30 --warmup 5 --no-cpu-baseline --no-ddp-leg --no-reference-gpu"
for shape in "$@"; do
  set -- $shape
  PD_STREAM_HS=$1 PD_STREAM_NST=$2 python bench.py $B > $O/${TAG}_hs$1_nst$2.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("$O/${TAG}_hs$1_nst$2.json")); print("hs=$1 nst=$2", "%.4f ms"%d["ms_per_step"], {k:round(x,4) for k,x in d["roofline"]["all_kernels_ms"].items()})
PY
done
