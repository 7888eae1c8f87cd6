#!/bin/bash
TAG=${1:-r2l}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/${TAG}_pytest.log
(cd scratch/r1tree && python bench.py --config cfg4 --steps 30 --warmup 5 --no-cpu-baseline > ../../$O/${TAG}_r1_cfg4.json 2>/dev/null)
python bench.py --config cfg4 --steps 30 --warmup 5 --no-cpu-baseline --no-ddp-leg --no-reference-gpu > $O/${TAG}_r2_cfg4.json 2>/dev/null
python bench.py --no-cpu-baseline --no-ddp-leg --no-reference-gpu > $O/${TAG}_r2_cfg2.json 2>/dev/null
python bench.py --config cfg3 --steps 30 --warmup 5 --no-cpu-baseline --no-ddp-leg --no-reference-gpu > $O/${TAG}_r2_cfg3.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:resize -c 4 --csv --log-file $O/${TAG}_resize_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-ddp-leg --no-reference-gpu > /dev/null 2>&1
grep resize $O/${TAG}_resize_launches.csv | awk -F'","' '{print $5, $NF}' | head -4
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2l_*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], "%.4f ms  %.0f img/s"%(d["ms_per_step"], d["value"]), {k:round(v,4) for k,v in d["roofline"]["all_kernels_ms"].items()}, "e2e", round(d["e2e"]["value"]), d["e2e"].get("other_transport",{}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
