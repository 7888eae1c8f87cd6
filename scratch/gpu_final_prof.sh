#!/bin/bash
# evidence set for profiles/: launch list of the default bench command + full captures of every library kernel
TAG=${1:-r1b}
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'rows_|ssim|photometric_bwd' -s 8 -c 4 -o $O/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $O/${TAG}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'rows_' -s 4 -c 2 -o $O/${TAG}c_prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --layout compact > $O/${TAG}c_ncu_full.log 2>&1
tail -2 $O/${TAG}_ncu_full.log | cut -c1-120
