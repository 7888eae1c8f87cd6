#!/bin/bash
# homography path: parity tests (oracle, goldens, full-size cfg4, the PR1 gate) + two cfg4 bench lines
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_trainer_gate.py -x -q -m gpu -k "oracle or golden or cfg4 or homography or gate" 2>&1 | tail -n 2
for i in 1 2; do
  python bench.py --config cfg4 --steps 30 --warmup 5 --no-cpu-baseline --no-ddp-leg --no-reference-gpu 2>/dev/null > gpurun_out/cfg4_$i.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/cfg4_$i.json").read().strip().splitlines()[-1]); print("cfg4 %.4f ms %.0f img/s"%(d["ms_per_step"], d["value"]), {k:round(v,4) for k,v in d["roofline"]["all_kernels_ms"].items()})
PY
done
