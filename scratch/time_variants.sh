#!/bin/bash
# usage: time_variants.sh CFG name1 name2 ...   -> per-kernel ms for each variant library
CFG=$1; shift
for v in "$@"; do
  PLANEDEPTH_B200_LIB=$PWD/scratch/variants/lib_$v.so python bench.py --config $CFG --steps 10 --warmup 3 --layout compact --no-cpu-baseline > /tmp/b_$v.json 2>/tmp/b_$v.err
  python - <<PY
import json
try:
    d=json.load(open("/tmp/b_$v.json")); k=d["roofline"]["all_kernels_ms"]
    print("%-16s step %.3f ms | fwd %.3f bwd %.3f lossf %.3f lossb %.3f" % ("$v", d["ms_per_step"], k["pd_warp_composite_fwd"], k["pd_warp_composite_bwd"], k["pd_photometric_fwd"], k["pd_photometric_bwd"]))
except Exception as e:
    print("$v FAILED", e); print(open("/tmp/b_$v.err").read()[-600:])
PY
done
