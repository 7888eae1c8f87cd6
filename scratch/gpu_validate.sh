#!/bin/bash
# Quick validation: parity tests + default bench + compact bench + smoke.
TAG=${1:-r1c}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err; echo "bench rc=$?"
python bench.py --steps 20 --warmup 5 --layout compact --no-cpu-baseline > $O/${TAG}_bench_cfg2_compact.json 2> $O/${TAG}_bench_cfg2_compact.err
tail -3 $O/${TAG}_pytest.log; tail -1 $O/${TAG}_smoke.log
cat $O/${TAG}_bench_cfg2.json
