#!/bin/bash
TAG=${1:-r2m}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/${TAG}_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:resize -c 4 --csv --log-file $O/${TAG}_resize_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-ddp-leg --no-reference-gpu > /dev/null 2>&1
grep resize $O/${TAG}_resize_launches.csv | awk -F'","' '{print $5, $NF}' | head -4
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:resize -c 2 --csv --log-file $O/${TAG}_resize_launches_cfg3.csv python bench.py --config cfg3 --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-ddp-leg --no-reference-gpu > /dev/null 2>&1
grep resize $O/${TAG}_resize_launches_cfg3.csv | awk -F'","' '{print $5, $NF}' | head -2
