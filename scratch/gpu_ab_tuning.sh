#!/bin/bash
# usage: gpu_ab_tuning.sh TAG "ENV1=.." "ENV2=.." ...  — cfg2 bench line per environment setting, on one box
TAG=$1; shift
O=gpurun_out
mkdir -p $O
B="python bench.py --no-cpu-baseline --steps 200 --warmup 10 --no-ddp-leg --no-reference-gpu"
i=0
$B > $O/${TAG}_base.json 2>/dev/null
for e in "$@"; do
  i=$((i+1))
  env $e $B > $O/${TAG}_v$i.json 2>/dev/null
  echo "v$i = $e"
done
$B > $O/${TAG}_base2.json 2>/dev/null
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], "%.4f ms"%d["ms_per_step"], {k:round(v,4) for k,v in d["roofline"]["all_kernels_ms"].items()}, "loss", d["loss"])
    except Exception as e: print(f, "ERR", e)
PY
