#!/bin/bash
TAG=${1:-r2z}
O=gpurun_out
mkdir -p $O
for cfg in "4 49 384 1280 1" "8 63 384 1280 1" "8 63 384 1280 0" "12 63 192 640 1"; do
  t=$(echo $cfg | tr ' ' '_')
  for mode in 0 1; do
    PD_TAIL_DIRECT=$mode ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tail_' -c 2 --csv --log-file $O/${TAG}_l_${t}_$mode.csv python scratch/tail_bench.py $cfg > /dev/null 2>&1
    echo "$cfg direct=$mode: $(grep tail_ $O/${TAG}_l_${t}_$mode.csv | awk -F'","' '{print $5, $NF}' | sed 's/(TailParams.*) / /; s/void //' | tr '\n' ' ' | cut -c1-200)"
  done
done
