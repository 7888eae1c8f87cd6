#!/bin/bash
# 2-GPU run: bench under torchrun with NCCL_DEBUG=INFO (stdout must stay ONE json line), ddp leg, e2e with copy-stream conversion
TAG=${1:-r2f}
O=gpurun_out
mkdir -p $O
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > $O/${TAG}_bench_2gpu.json 2> $O/${TAG}_bench_2gpu.err; echo "2gpu rc=$?"
wc -l $O/${TAG}_bench_2gpu.json; grep -c "NCCL INFO" $O/${TAG}_bench_2gpu.err; grep -m3 "comm 0x.*rank .* nranks" $O/${TAG}_bench_2gpu.err
echo skip-1gpu
python - <<'PY'
import json
for f in ["gpurun_out/r2f_bench_2gpu.json"]:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["n_gpus"], "%.4f ms  %.0f img/s"%(d["ms_per_step"], d["value"]))
    print("   e2e", d["e2e"])
    print("   ddp", d.get("ddp"))
PY
