# end-of-round refresh: default bench line, then the ncu launch list of the resident step
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r1_bench_default.json 2> gpurun_out/r1_bench_default.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e"
$CMD > gpurun_out/r1_plain.json 2> gpurun_out/r1_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1_launches_wgs30x.csv $CMD > gpurun_out/ncu_ll.log 2>&1
tail -c 900 gpurun_out/r1_bench_default.json
