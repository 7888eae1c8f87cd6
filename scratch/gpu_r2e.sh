#!/bin/bash
TAG=${1:-r2e}
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_perceptual.py tests/test_input_staging.py -m gpu -q > $O/${TAG}_pytest_new.log 2>&1; echo "pytest new rc=$?"; tail -4 $O/${TAG}_pytest_new.log
python bench.py --no-cpu-baseline > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err; echo "bench rc=$?"; tail -3 $O/${TAG}_bench_cfg2.err
python bench.py --no-cpu-baseline --config cfg3 --steps 30 --no-ddp-leg --no-reference-gpu > $O/${TAG}_bench_cfg3.json 2> $O/${TAG}_bench_cfg3.err; echo "cfg3 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/${TAG}_launches_fused.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-ddp-leg --no-reference-gpu > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/${TAG}_launches_unfused.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-fuse-bwd --no-ddp-leg --no-reference-gpu > /dev/null 2>&1
python - <<'PY'
import json,glob,csv
for f in sorted(glob.glob("gpurun_out/r2e_bench*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], "%.4f ms  %.0f img/s"%(d["ms_per_step"], d["value"]), {k:round(v,4) for k,v in d["roofline"]["all_kernels_ms"].items()}, "frac", round(d["roofline"]["frac"],3))
        print("    e2e", d["e2e"])
        for k in ("ddp","reference_gpu"):
            if k in d: print("   ", k, d[k])
    except Exception as e: print(f, "ERR", e)
for t in ["fused","unfused"]:
    rows=list(csv.reader(open("gpurun_out/r2e_launches_%s.csv"%t)))
    for i,r in enumerate(rows):
        if "Kernel Name" in r: hdr=r; start=i+1; break
    ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
    seq=[(r[ki][:70], r[vi]) for r in rows[start:] if len(r)>vi]
    print(t)
    for k,v in seq[-8:]: print("   ",k,v)
PY
