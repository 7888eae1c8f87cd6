#!/bin/bash
TAG=${1:-b}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${TAG}_pytest.log
for c in cfg4 cfg5 cfg3; do for l in reference compact; do
  timeout 600 python bench.py --config $c --layout $l --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_${c}_$l.json 2> $O/${TAG}_${c}_$l.err
done; done
python - <<PY
import json
for c in ["cfg3","cfg4","cfg5"]:
  for l in ["reference","compact"]:
    try:
        d=json.load(open("$O/${TAG}_%s_%s.json"%(c,l))); k=d["roofline"]
        print("%-5s %-9s %.0f img/s (e2e %.0f) step %.3f ms | dom %s %.3f ms frac %.3f | %s" % (c, l, d["value"], d["e2e"]["value"], d["ms_per_step"], k["kernel"], k["kernel_ms"], k["frac"], {a: round(b,3) for a,b in k["all_kernels_ms"].items()}))
    except Exception as e:
        print(c, l, "FAILED", e); print(open("$O/${TAG}_%s_%s.err"%(c,l)).read()[-500:])
PY
