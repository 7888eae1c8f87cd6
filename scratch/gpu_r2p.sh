#!/bin/bash
TAG=${1:-r2p}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bf16" > $O/${TAG}_pytest_bf16.log 2>&1; echo "bf16 tests rc=$?"; grep -E "^E  |passed|failed|FAILED" $O/${TAG}_pytest_bf16.log | head -30 | cut -c1-250
