#!/bin/bash
# compute-sanitizer over a slice of the parity suite (small shapes: every kernel family, bf16 storage, promise check,
# both decoder-tail kernel families): memcheck, then racecheck on the kernels that use shared memory + mbarriers
TAG=${1:-r2u}
O=gpurun_out
mkdir -p $O
SEL="None-0- or None-2- or ssim_l1-4- or None-7- or None-10- or None-14- or None-17- or bf16_storage_kernels or promise or post_process_disp_matches_oracle or decoder_tail_matches_oracle or smooth or resize or persistent"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_input_staging.py tests/test_gpu_fullsize.py -m gpu -q -x -k "$SEL" > $O/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $O/${TAG}_memcheck.log | head -12
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "None-0- or ssim_l1-4- or None-10- or decoder_tail_matches_oracle" > $O/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed|hazard" $O/${TAG}_racecheck.log | head -12
