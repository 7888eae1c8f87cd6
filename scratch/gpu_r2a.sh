#!/bin/bash
# Round 2, call A: existing parity suite on the refactored library, the new full-size / persistent-loop / promise / PR1-gate
# tests in REPORT mode (error statistics instead of assertions), smoke, one bench line.
TAG=${1:-r2a}
O=gpurun_out
mkdir -p $O
nproc > $O/${TAG}_host.txt; free -g >> $O/${TAG}_host.txt
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/${TAG}_pytest_parity.log 2>&1; echo "parity rc=$?"; tail -3 $O/${TAG}_pytest_parity.log
PD_TEST_REPORT=${TAG}_fullsize_report.json python -m pytest tests/test_gpu_fullsize.py -m gpu -q --durations=0 > $O/${TAG}_pytest_fullsize.log 2>&1; echo "fullsize rc=$?"; tail -15 $O/${TAG}_pytest_fullsize.log
PD_TEST_REPORT=1 python -m pytest tests/test_gpu_trainer_gate.py -m gpu -q -s > $O/${TAG}_pytest_gate.log 2>&1; echo "gate rc=$?"; tail -12 $O/${TAG}_pytest_gate.log
python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/${TAG}_smoke.log
python bench.py --no-cpu-baseline > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err; echo "bench rc=$?"
cat $O/${TAG}_bench_cfg2.json
