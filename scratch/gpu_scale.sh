#!/bin/bash
# usage: gpu_scale.sh N TAG  — the driver's own launch line for N GPUs
N=$1; TAG=${2:-r2s}
O=gpurun_out
mkdir -p $O
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 200 --warmup 10 > $O/${TAG}_bench_${N}gpu.json 2> $O/${TAG}_bench_${N}gpu.err; echo "rc=$?"
wc -l $O/${TAG}_bench_${N}gpu.json
grep -c "NCCL INFO" $O/${TAG}_bench_${N}gpu.err
grep -m2 "NVLS\|via P2P\|Connected all rings\|Channel 00/" $O/${TAG}_bench_${N}gpu.err | cut -c1-200
grep -v "NCCL INFO" $O/${TAG}_bench_${N}gpu.err | tail -n 5
python - <<PY
import json
d=json.loads(open("$O/${TAG}_bench_${N}gpu.json").read().strip().splitlines()[-1])
print(d["n_gpus"], "%.4f ms  %.0f img/s"%(d["ms_per_step"], d["value"]))
print("   e2e", d["e2e"])
for k in ("ddp","cfg4_dp4","cfg5_dp8"):
    if k in d: print("  ", k, d[k])
PY
# keep the stderr small enough to bring back
head -c 200000 $O/${TAG}_bench_${N}gpu.err > $O/${TAG}_bench_${N}gpu.err.head; mv $O/${TAG}_bench_${N}gpu.err.head $O/${TAG}_bench_${N}gpu.err
