#!/bin/bash
# bf16 storage: parity tests + bench A/B against fp32 storage (cfg2, cfg3)
TAG=${1:-r2b}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu -k "bf16" > $O/${TAG}_pytest.log 2>&1; echo "bf16 tests rc=$?"; tail -n 3 $O/${TAG}_pytest.log
B="--steps 40 --warmup 5 --no-cpu-baseline --no-ddp-leg --no-reference-gpu"
for cfg in cfg2 cfg3; do
for st in fp32 bf16 fp32 bf16; do
  python bench.py --config $cfg --storage $st $B > $O/${TAG}_${cfg}_$st.json 2>$O/${TAG}_err.log || tail -n 3 $O/${TAG}_err.log
  python - <<PY
import json
d=json.load(open("$O/${TAG}_${cfg}_$st.json")); print("$cfg $st", "%.4f ms"%d["ms_per_step"], {k:round(x,4) for k,x in d["roofline"]["all_kernels_ms"].items()})
PY
done
done
