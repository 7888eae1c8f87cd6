#!/bin/bash
TAG=${1:-r2r}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_bench_gpu.py -m gpu -q > $O/${TAG}_pytest_bench.log 2>&1; echo "bench test rc=$?"; tail -n 5 $O/${TAG}_pytest_bench.log | cut -c1-300
python bench.py > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err; echo "bench rc=$?"; tail -n 3 $O/${TAG}_bench_cfg2.err | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; echo "ref arm rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2r_bench_cfg2.json"))
print("%.4f ms  %.0f img/s"%(d["ms_per_step"], d["value"]), {k:round(v,4) for k,v in d["roofline"]["all_kernels_ms"].items()}, "frac", round(d["roofline"]["frac"],3), "step_frac", round(d["roofline"]["step_frac"],3))
print("e2e", d["e2e"]["value"], d["e2e"]["other_transport"]["value"]); print("cpu", d["cpu_baseline"]); print("ddp", d["ddp"]["value"]); print("refgpu", d["reference_gpu"]["value"], d["clocks"])
print(open("gpurun_out/r2r_bench_reference.json").read()[:500])
PY
