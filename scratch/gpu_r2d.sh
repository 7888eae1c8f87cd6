#!/bin/bash
# Round 2, call D: GPU suite on the micro-optimised stream kernels + fused backward without grad materialisation, bench lines.
TAG=${1:-r2d}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $O/${TAG}_pytest.log
python bench.py --no-cpu-baseline > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err; echo "bench rc=$?"; tail -3 $O/${TAG}_bench_cfg2.err
python bench.py --no-cpu-baseline --no-fuse-bwd --no-ddp-leg --no-reference-gpu --steps 100 > $O/${TAG}_bench_cfg2_unfused.json 2>/dev/null
for c in cfg3 cfg4 cfg5; do
  python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline --no-ddp-leg --no-reference-gpu > $O/${TAG}_bench_$c.json 2> $O/${TAG}_bench_$c.err; echo "$c rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2d_bench*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], "%.4f ms  %.0f img/s  e2e %.0f"%(d["ms_per_step"], d["value"], d["e2e"]["value"]), {k:round(v,4) for k,v in d["roofline"]["all_kernels_ms"].items()}, "frac", round(d["roofline"]["frac"],3))
        for k in ("ddp","reference_gpu"):
            if k in d: print("   ", k, d[k])
    except Exception as e: print(f, "ERR", e)
PY
