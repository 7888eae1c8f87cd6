#!/bin/bash
# Round 2, call J: cfg4 regression diagnosis (fused / unfused, ncu kernel list) + profile set r2 (launch list, full capture).
TAG=${1:-r2j}
O=gpurun_out
mkdir -p $O
B4="python bench.py --config cfg4 --steps 30 --warmup 5 --no-cpu-baseline --no-ddp-leg --no-reference-gpu"
$B4 > $O/${TAG}_cfg4_fused.json 2>/dev/null
$B4 --no-fuse-bwd > $O/${TAG}_cfg4_unfused.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${TAG}_cfg4_launches.csv $B4 --steps 2 --warmup 3 --no-graph > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${TAG}_cfg4_launches_unfused.csv $B4 --steps 2 --warmup 3 --no-graph --no-fuse-bwd > /dev/null 2>&1
# profile set of the headline config
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ddp-leg --no-reference-gpu > $O/r2_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'rows_|ssim_l1_stream' -s 12 -c 3 -o $O/r2_prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-ddp-leg --no-reference-gpu > $O/r2_ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:'rows_' -s 4 -c 2 -o $O/r2_prof_cfg3 -f python bench.py --config cfg3 --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-ddp-leg --no-reference-gpu > $O/r2_ncu_cfg3.log 2>&1
ncu --set full --clock-control none -k regex:'homo_' -s 6 -c 2 -o $O/r2_prof_cfg4 -f python bench.py --config cfg4 --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-ddp-leg --no-reference-gpu > $O/r2_ncu_cfg4.log 2>&1
ncu --set full --clock-control none -k regex:'resize_bicubic' -c 1 -o $O/r2_prof_resize -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-ddp-leg --no-reference-gpu > $O/r2_ncu_resize.log 2>&1
for r in r2_prof r2_prof_cfg3 r2_prof_cfg4 r2_prof_resize; do
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
done
ncu -i $O/r2_prof.ncu-rep --page source --csv --kernel-name regex:rows_bwd > $O/r2_src_rows_bwd.csv 2>/dev/null
ncu -i $O/r2_prof.ncu-rep --page source --csv --kernel-name regex:rows_fwd > $O/r2_src_rows_fwd.csv 2>/dev/null
ncu -i $O/r2_prof.ncu-rep --page source --csv --kernel-name regex:ssim_l1_stream > $O/r2_src_ssim.csv 2>/dev/null
rm -f $O/r2_prof_cfg3.ncu-rep $O/r2_prof_cfg4.ncu-rep $O/r2_prof_resize.ncu-rep
python - <<'PY'
import json,csv
for t in ["fused","unfused"]:
    d=json.load(open("gpurun_out/r2j_cfg4_%s.json"%t)); print(t, "%.4f ms"%d["ms_per_step"], {k:round(v,4) for k,v in d["roofline"]["all_kernels_ms"].items()})
for t in ["","_unfused"]:
    rows=list(csv.reader(open("gpurun_out/r2j_cfg4_launches%s.csv"%t)))
    for i,r in enumerate(rows):
        if "Kernel Name" in r: hdr=r; start=i+1; break
    ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
    seq=[(r[ki][:60], r[vi]) for r in rows[start:] if len(r)>vi]
    print("launch list", t or "fused")
    for k,v in seq[-26:]: print("   ",k,v)
PY
du -sh $O
