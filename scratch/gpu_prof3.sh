#!/bin/bash
# ncu full capture (with source) of the step's kernels: streamed warp fwd/bwd + SSIM stream; usage: gpu_prof3.sh TAG [bench args]
TAG=${1:-p}; shift
O=gpurun_out
mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:'rows_|ssim_l1_stream|warp_composite' -s 9 -c 3 -o $O/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph "$@" > $O/${TAG}_ncu_full.log 2>&1
tail -3 $O/${TAG}_ncu_full.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > $O/${TAG}_launches.log 2>&1
ls -la $O | tail -5
