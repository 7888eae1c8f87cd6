#!/bin/bash
# ncu full capture of the homography kernels (cfg4) + raw CSV for profiles/summarize.py-style summaries
O=gpurun_out
mkdir -p $O
ncu --set full --clock-control none -k regex:'homo_' -s 6 -c 2 -o $O/r2_prof_cfg4 -f python bench.py --config cfg4 --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-ddp-leg --no-reference-gpu > $O/r2_ncu_cfg4.log 2>&1
ncu -i $O/r2_prof_cfg4.ncu-rep --page raw --csv > $O/r2_prof_cfg4_raw.csv 2>/dev/null
rm -f $O/r2_prof_cfg4.ncu-rep
ls -la $O/r2_prof_cfg4_raw.csv
