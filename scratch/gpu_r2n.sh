#!/bin/bash
TAG=${1:-r2n}
O=gpurun_out
mkdir -p $O
python scratch/prof_step.py cfg5 reference > $O/${TAG}_prof_cfg5.txt 2>&1
python bench.py --config cfg5 --steps 20 --warmup 5 --no-cpu-baseline --no-ddp-leg --no-reference-gpu > $O/${TAG}_bench_cfg5.json 2>/dev/null
python bench.py --config cfg5 --layout compact --steps 20 --warmup 5 --no-cpu-baseline --no-ddp-leg --no-reference-gpu > $O/${TAG}_bench_cfg5_compact.json 2>/dev/null
grep -v "^---" $O/${TAG}_prof_cfg5.txt | cut -c1-75,150-260 | head -30
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2n_*.json")):
    d=json.load(open(f)); print(f.split("/")[-1], "%.4f ms  %.0f img/s"%(d["ms_per_step"], d["value"]), {k:round(v,4) for k,v in d["roofline"]["all_kernels_ms"].items()})
PY
