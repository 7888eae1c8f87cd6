#!/bin/bash
# Round 2, call C: fused vs unfused photometric backward (graph timings + ncu kernel durations), ring-shape sweep.
TAG=${1:-r2c}
O=gpurun_out
mkdir -p $O
B="python bench.py --no-cpu-baseline --steps 100 --warmup 5"
$B > $O/${TAG}_fused.json 2>/dev/null
$B --no-fuse-bwd > $O/${TAG}_unfused.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${TAG}_launches_fused.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${TAG}_launches_unfused.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-fuse-bwd > /dev/null 2>&1
for hs in 4 6 8; do for nst in 2 3 4; do
  PD_STREAM_HS=$hs PD_STREAM_NST=$nst $B --no-fuse-bwd > $O/${TAG}_cfg2_hs${hs}_nst${nst}.json 2>/dev/null
  PD_STREAM_HS=$hs PD_STREAM_NST=$nst $B --no-fuse-bwd --config cfg3 --steps 30 > $O/${TAG}_cfg3_hs${hs}_nst${nst}.json 2>/dev/null
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c_*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], "%.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["roofline"]["all_kernels_ms"].items()})
    except Exception as e: print(f, "ERR", e)
PY
