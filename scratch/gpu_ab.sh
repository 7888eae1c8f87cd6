#!/bin/bash
# parity tests + smoke + A/B benches of the switches given as "NAME=VAL" env settings
# usage: scratch/gpu_ab.sh TAG "ENV1=1" "ENV2=1|--extra-bench-args" ...
TAG=${1:-ab}; shift
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest.log
python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/${TAG}_smoke.log
python bench.py --steps 50 --warmup 5 > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err; echo "bench rc=$?"
python bench.py --steps 50 --warmup 5 --layout compact --no-cpu-baseline > $O/${TAG}_bench_cfg2_compact.json 2> $O/${TAG}_bench_cfg2_compact.err
i=0
for e in "$@"; do
  i=$((i+1))
  envs="${e%%|*}"; extra=""; [[ "$e" == *"|"* ]] && extra="${e#*|}"
  env $envs python bench.py --steps 50 --warmup 5 --no-cpu-baseline $extra > $O/${TAG}_bench_cfg2_alt$i.json 2> $O/${TAG}_bench_cfg2_alt$i.err; echo "alt$i ($e) rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/${TAG}_bench_*.json")):
    try:
        j=json.load(open(f)); print(f, round(j["value"]), {k:round(v,4) for k,v in j["roofline"]["all_kernels_ms"].items()})
    except Exception as e: print(f, e)
PY
