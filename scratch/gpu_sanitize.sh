#!/bin/bash
# compute-sanitizer memcheck over a slice of the parity suite (small shapes: every kernel family, bf16 storage, promise check)
TAG=${1:-r2u}
O=gpurun_out
mkdir -p $O
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_input_staging.py -m gpu -q -x -k "None-0- or None-2- or ssim_l1-4- or None-7- or None-10- or None-14- or None-17- or bf16_storage_kernels or promise or post_process_disp_matches_oracle or decoder_tail_matches_oracle or smooth or resize" > $O/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $O/${TAG}_memcheck.log | head -12
timeout 900 python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 $O/${TAG}_pytest.log
python bench.py --no-cpu-baseline --no-ddp-leg --no-reference-gpu > $O/${TAG}_bench_cfg2.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_cfg2.json')); print('%.4f ms'%d['ms_per_step'], {k:round(v,4) for k,v in d['roofline']['all_kernels_ms'].items()}, 'frac', round(d['roofline']['frac'],3))"
