#!/bin/bash
# Round 2 validation set: full GPU suite, smoke, bench default (all legs), refreshed profile set for the final kernels.
TAG=${1:-r2q}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 $O/${TAG}_pytest.log | cut -c1-200
python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 $O/${TAG}_smoke.log
python bench.py > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err; echo "bench rc=$?"
for c in cfg3 cfg4 cfg5; do python bench.py --config $c --no-cpu-baseline --no-ddp-leg --no-reference-gpu > $O/${TAG}_bench_$c.json 2>/dev/null; done
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; echo "ref arm rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ddp-leg --no-reference-gpu > $O/r2_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'rows_|ssim_l1_stream' -s 12 -c 3 -o $O/r2_prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-ddp-leg --no-reference-gpu > $O/r2_ncu_full.log 2>&1
ncu -i $O/r2_prof.ncu-rep --page raw --csv > $O/r2_prof_raw.csv 2>/dev/null
ncu -i $O/r2_prof.ncu-rep --page source --csv --kernel-name regex:rows_bwd > $O/r2_src_rows_bwd.csv 2>/dev/null
ncu -i $O/r2_prof.ncu-rep --page source --csv --kernel-name regex:rows_fwd > $O/r2_src_rows_fwd.csv 2>/dev/null
ncu -i $O/r2_prof.ncu-rep --page source --csv --kernel-name regex:ssim_l1_stream > $O/r2_src_ssim.csv 2>/dev/null
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2q_bench_cfg2.json"))
print("%.4f ms  %.0f img/s"%(d["ms_per_step"], d["value"]), {k:round(v,4) for k,v in d["roofline"]["all_kernels_ms"].items()}, "frac", round(d["roofline"]["frac"],3), "step_frac", round(d["roofline"]["step_frac"],3))
print("e2e", d["e2e"]["value"], d["e2e"]["other_transport"]["value"]); print("cpu", d["cpu_baseline"]); print("ddp", d["ddp"]["value"]); print("refgpu", d["reference_gpu"]["value"], d["clocks"])
print(open("gpurun_out/r2q_bench_reference.json").read()[:400])
for c in ['cfg3','cfg4','cfg5']:
    d=json.load(open('gpurun_out/r2q_bench_%s.json'%c)); print(c, '%.4f ms  %.0f img/s'%(d['ms_per_step'], d['value']), {k:round(v,4) for k,v in d['roofline']['all_kernels_ms'].items()}, 'e2e', round(d['e2e']['value']))
PY
