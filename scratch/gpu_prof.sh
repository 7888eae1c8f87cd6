#!/bin/bash
# ncu full capture of the warp kernels (one launch each) + source-level hot spots; usage: gpu_prof.sh TAG [bench args]
TAG=${1:-p}; shift
O=gpurun_out
mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:'rows_|warp_composite' -s 6 -c 2 -o $O/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph "$@" > $O/${TAG}_ncu_full.log 2>&1
tail -3 $O/${TAG}_ncu_full.log | cut -c1-200
