#!/bin/bash
# bench lines for every config + 2-rank sanity; usage: gpu_bench.sh TAG
TAG=${1:-b}
O=gpurun_out
mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/${TAG}_cfg2.json 2> $O/${TAG}_cfg2.err; echo "cfg2 rc=$?"
for c in cfg3 cfg4 cfg5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_$c.json 2> $O/${TAG}_$c.err; echo "$c rc=$?"
done
python bench.py --steps 20 --warmup 5 --layout compact --no-cpu-baseline > $O/${TAG}_cfg2_compact.json 2> $O/${TAG}_cfg2_compact.err
python - <<PY
import json
for c in ["cfg2","cfg2_compact","cfg3","cfg4","cfg5"]:
    try:
        d=json.load(open("$O/${TAG}_%s.json"%c)); k=d["roofline"]
        print("%-13s %.0f img/s (e2e %.0f) step %.3f ms | dom %s %.3f ms frac %.3f | %s" % (c, d["value"], d["e2e"]["value"], d["ms_per_step"], k["kernel"], k["kernel_ms"], k["frac"], {a: round(b,3) for a,b in k["all_kernels_ms"].items()}))
    except Exception as e:
        print(c, "FAILED", e); print(open("$O/${TAG}_%s.err"%c).read()[-600:])
PY
