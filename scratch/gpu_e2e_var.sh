#!/bin/bash
# run-to-run spread of the e2e leg (both transports are measured in every run; the headline is the faster one)
for cfg in cfg2 cfg2 cfg3; do
  python bench.py --config $cfg --no-cpu-baseline --no-ddp-leg --no-reference-gpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; o=e['other_transport']; print('$cfg e2e %.0f img/s via %s (%.1f GB/s); other: %s %.0f; value %.0f' % (e['value'], e['colour_transport'], e['h2d_GBps'], o['colour_transport'], o['value'], d['value']))"
done
timeout 300 python -m pytest tests/test_bench_gpu.py -x -q -m gpu 2>&1 | tail -n 2
