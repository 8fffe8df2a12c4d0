set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e"
$CMD > gpurun_out/ll_plain.json 2> gpurun_out/ll_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_wgs.csv $CMD > gpurun_out/ncu_ll.log 2>&1
tail -2 gpurun_out/ncu_ll.log
