#!/bin/bash
# Round 2, call B: whole GPU suite asserting (fused photometric backward on), full-size report, smoke, bench.
TAG=${1:-r2b}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 $O/${TAG}_pytest.log
PD_TEST_REPORT=${TAG}_fullsize_report.json python -m pytest tests/test_gpu_fullsize.py -m gpu -q > $O/${TAG}_pytest_fullsize_report.log 2>&1; echo "fullsize report rc=$?"
python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/${TAG}_smoke.log
python bench.py --no-cpu-baseline > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err; echo "bench rc=$?"
cat $O/${TAG}_bench_cfg2.json
for c in cfg3 cfg4 cfg5; do
  python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_$c.json 2> $O/${TAG}_bench_$c.err; echo "$c rc=$?"
done
