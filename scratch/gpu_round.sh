#!/bin/bash
# One GPU session: parity tests, bench (both layouts, cfg3, reference arm), ncu launch list + full capture.
# usage: scratch/gpu_round.sh TAG
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" 
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err; echo "bench rc=$?"
python bench.py --steps 20 --warmup 5 --layout compact --no-cpu-baseline > $O/${TAG}_bench_cfg2_compact.json 2> $O/${TAG}_bench_cfg2_compact.err
python bench.py --config cfg3 --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_cfg3.json 2> $O/${TAG}_bench_cfg3.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:warp_composite -s 6 -c 2 -o $O/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $O/${TAG}_ncu_full.log 2>&1
[ -x scratch/ffma2_bench ] && scratch/ffma2_bench > $O/${TAG}_ffma2.log 2>&1
tail -3 $O/${TAG}_pytest.log
cat $O/${TAG}_bench_cfg2.json
