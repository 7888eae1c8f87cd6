#!/bin/bash
# Round-end measurement set: parity + smoke, bench lines of every config (+ promise-free / compact variants), reference arms
# (CPU arm of bench.py, GPU probe of the unmodified reference), ncu launch list + full capture of the step's kernels.
TAG=${1:-r1z}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${TAG}_pytest.log
python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err; echo "bench rc=$?"
python bench.py --no-rowwise --no-cpu-baseline > $O/${TAG}_bench_cfg2_nopromise.json 2> $O/${TAG}_bench_cfg2_nopromise.err
python bench.py --layout compact --no-cpu-baseline > $O/${TAG}_bench_cfg2_compact.json 2> $O/${TAG}_bench_cfg2_compact.err
for c in cfg3 cfg4 cfg5; do
  python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_$c.json 2> $O/${TAG}_bench_$c.err; echo "$c rc=$?"
done
python bench.py --config cfg5 --layout compact --steps 30 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_cfg5_compact.json 2> $O/${TAG}_bench_cfg5_compact.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; echo "ref arm rc=$?"
python scratch/ref_gpu_probe.py cfg2 cfg3 cfg4 cfg5 > $O/${TAG}_reference_gpu_probe.json 2> $O/${TAG}_reference_gpu_probe.err; echo "ref gpu probe rc=$?"
python scratch/pp_bench.py 12 49 192 640 > $O/${TAG}_pp_cfg2.json 2> $O/${TAG}_pp.err
python scratch/pp_bench.py 8 63 384 1280 > $O/${TAG}_pp_cfg5.json 2>> $O/${TAG}_pp.err
python scratch/tail_bench.py 12 49 192 640 0 > $O/${TAG}_tail_cfg2.json 2> $O/${TAG}_tail.err
python scratch/tail_bench.py 4 49 384 1280 1 > $O/${TAG}_tail_cfg3.json 2>> $O/${TAG}_tail.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'rows_|ssim_l1_stream|photometric_bwd' -s 12 -c 4 -o $O/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $O/${TAG}_ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:'rows_|homo_' -s 4 -c 2 -o $O/${TAG}_prof_cfg3 -f python bench.py --config cfg3 --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $O/${TAG}_ncu_cfg3.log 2>&1
ncu --set full --clock-control none -k regex:'homo_' -s 6 -c 2 -o $O/${TAG}_prof_cfg4 -f python bench.py --config cfg4 --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $O/${TAG}_ncu_cfg4.log 2>&1
# the reports are large (gpurun brings back <= 64 MiB): keep CSV exports, drop the two secondary reports
for r in prof prof_cfg3 prof_cfg4; do
  ncu -i $O/${TAG}_$r.ncu-rep --page raw --csv > $O/${TAG}_${r}_raw.csv 2>/dev/null
done
ncu -i $O/${TAG}_prof.ncu-rep --page source --csv --kernel-name regex:rows_bwd > $O/${TAG}_src_rows_bwd.csv 2>/dev/null
ncu -i $O/${TAG}_prof.ncu-rep --page source --csv --kernel-name regex:rows_fwd > $O/${TAG}_src_rows_fwd.csv 2>/dev/null
ncu -i $O/${TAG}_prof.ncu-rep --page source --csv --kernel-name regex:ssim_l1_stream > $O/${TAG}_src_ssim.csv 2>/dev/null
rm -f $O/${TAG}_prof_cfg3.ncu-rep $O/${TAG}_prof_cfg4.ncu-rep
du -sh $O
cat $O/${TAG}_bench_cfg2.json
ls $O | grep $TAG | wc -l
