#!/bin/bash
TAG=${1:-r2w}
O=gpurun_out
mkdir -p $O
ncu --set full --clock-control none -k regex:'tail_' -c 4 -o $O/${TAG}_tail -f python scratch/tail_bench.py 12 49 192 640 0 > /dev/null 2>&1
ncu -i $O/${TAG}_tail.ncu-rep --page raw --csv > $O/${TAG}_tail_raw.csv 2>/dev/null
PD_TAIL_DIRECT=1 ncu --set full --clock-control none -k regex:'tail_' -c 4 -o $O/${TAG}_taild -f python scratch/tail_bench.py 12 49 192 640 0 > /dev/null 2>&1
ncu -i $O/${TAG}_taild.ncu-rep --page raw --csv > $O/${TAG}_taild_raw.csv 2>/dev/null
rm -f $O/${TAG}_tail.ncu-rep $O/${TAG}_taild.ncu-rep
python - <<'PY'
import csv
for f in ["gpurun_out/r2w_tail_raw.csv","gpurun_out/r2w_taild_raw.csv"]:
    rows=list(csv.reader(open(f))); hdr=rows[0]
    for r in rows[2:]:
        d=dict(zip(hdr,r))
        print(d["Kernel Name"][:40], "t", d.get("gpu__time_duration.sum"), "regs", d.get("launch__registers_per_thread"), "occ", d.get("sm__warps_active.avg.pct_of_peak_sustained_active"), "issue", d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"), "inst", d.get("smsp__inst_executed.sum"), "dramR", d.get("dram__bytes_read.sum"), "dramW", d.get("dram__bytes_write.sum"), "grid", d.get("launch__grid_size"), "smem", d.get("launch__shared_mem_per_block_dynamic"), "lim smem", d.get("launch__occupancy_limit_shared_mem"), "lim regs", d.get("launch__occupancy_limit_registers"))
PY
