#!/bin/bash
# parity tests + bench of one config with optional alternates; usage: gpu_cfg.sh TAG CFG "ENV|args" ...
TAG=$1; CFG=$2; shift; shift
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/${TAG}_pytest.log | cut -c1-300
python bench.py --config $CFG --steps 20 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_${CFG}.json 2> $O/${TAG}_bench_${CFG}.err; echo "bench rc=$?"; tail -3 $O/${TAG}_bench_${CFG}.err
i=0
for e in "$@"; do
  i=$((i+1))
  envs="${e%%|*}"; extra=""; [[ "$e" == *"|"* ]] && extra="${e#*|}"
  env $envs python bench.py --config $CFG --steps 20 --warmup 5 --no-cpu-baseline $extra > $O/${TAG}_bench_${CFG}_alt$i.json 2> $O/${TAG}_bench_${CFG}_alt$i.err; echo "alt$i ($e) rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/${TAG}_bench_*.json")):
    try:
        j=json.load(open(f)); print(f, round(j["value"]), {k:round(v,4) for k,v in j["roofline"]["all_kernels_ms"].items()})
    except Exception as e: print(f, e)
PY
