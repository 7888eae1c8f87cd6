#!/bin/bash
# quick iteration: parity tests under a timeout (a pipeline deadlock must not hang the box), then kernel timings
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $O/${TAG}_pytest.log
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --config ${CFG:-cfg2} --steps 10 --warmup 3 --layout ${LAYOUT:-compact} --no-cpu-baseline > $O/${TAG}_$name.json 2> $O/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$O/${TAG}_$name.json")); k=d["roofline"]["all_kernels_ms"]
    print("%-22s step %.3f ms %.0f img/s | fwd %.3f bwd %.3f lossf %.3f lossb %.3f" % ("$name", d["ms_per_step"], d["value"], k["pd_warp_composite_fwd"], k["pd_warp_composite_bwd"], k["pd_photometric_fwd"], k["pd_photometric_bwd"]))
except Exception as e:
    print("$name FAILED", e); print(open("$O/${TAG}_$name.err").read()[-800:])
PY
}
if [ -f scratch/variants.txt ]; then
  while read -r line; do [ -z "$line" ] && continue; eval "run $line"; done < scratch/variants.txt
fi
