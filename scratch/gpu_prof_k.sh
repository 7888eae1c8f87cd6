#!/bin/bash
# ncu full capture of kernels matching a regex in one bench config; usage: gpu_prof_k.sh TAG CFG REGEX [skip] [count]
TAG=$1; CFG=$2; RE=$3; SK=${4:-4}; CN=${5:-2}
O=gpurun_out
mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SK -c $CN -o $O/${TAG}_prof -f python bench.py --config $CFG --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $O/${TAG}_ncu_full.log 2>&1
tail -3 $O/${TAG}_ncu_full.log | cut -c1-200
