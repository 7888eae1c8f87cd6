#!/bin/bash
TAG=${1:-r2b2}
O=gpurun_out
mkdir -p $O
B="--steps 30 --warmup 5 --no-cpu-baseline --no-ddp-leg --no-reference-gpu --storage bf16"
for shape in "0 0" "4 3" "4 4" "5 3" "6 3" "8 2" "8 3" "6 2" "4 5"; do
  set -- $shape
  PD_STREAM_HS=$1 PD_STREAM_NST=$2 python bench.py $B > $O/${TAG}_hs$1_nst$2.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("$O/${TAG}_hs$1_nst$2.json")); print("bf16 hs=$1 nst=$2", "%.4f ms"%d["ms_per_step"], {k:round(x,4) for k,x in d["roofline"]["all_kernels_ms"].items()})
PY
done
