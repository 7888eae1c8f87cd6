#!/bin/bash
TAG=${1:-p}; shift
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > $O/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ssim|photometric|reduce_partials' -s 6 -c 3 -o $O/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph "$@" > $O/${TAG}_ncu_full.log 2>&1
tail -2 $O/${TAG}_ncu_full.log | cut -c1-150
