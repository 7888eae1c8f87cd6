#!/bin/bash
TAG=${1:-r2v}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "decoder_tail" > $O/${TAG}_pytest_tail.log 2>&1; echo "tail tests rc=$?"; grep -E "^E  |passed|failed|FAILED" $O/${TAG}_pytest_tail.log | head -12 | cut -c1-250
for cfg in "12 49 192 640 0" "4 49 384 1280 1" "8 63 384 1280 1"; do
  t=$(echo $cfg | tr ' ' '_')
  python scratch/tail_bench.py $cfg > $O/${TAG}_tail_tile_$t.json 2>/dev/null
  PD_TAIL_DIRECT=1 python scratch/tail_bench.py $cfg > $O/${TAG}_tail_direct_$t.json 2>/dev/null
  echo "$cfg tile:   $(cat $O/${TAG}_tail_tile_$t.json | cut -c1-200)"
  echo "$cfg direct: $(cat $O/${TAG}_tail_direct_$t.json | cut -c1-200)"
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tail_' -c 4 --csv --log-file $O/${TAG}_tail_launches.csv python scratch/tail_bench.py 12 49 192 640 0 > /dev/null 2>&1
grep tail_ $O/${TAG}_tail_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120
