# the measurements DESIGN.md / profiles/ quote: default bench, reference arm, launch list, ncu --set full of the top kernels
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r1_bench_default.json 2> gpurun_out/r1_bench_default.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r1_bench_reference.json 2> gpurun_out/r1_bench_reference.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e"
$CMD > gpurun_out/r1_plain.json 2> gpurun_out/r1_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1_launches_wgs30x.csv $CMD > gpurun_out/ncu_ll.log 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_depth_tiles16|k_walk|k_span_agg' -s 6 -c 3 -f -o gpurun_out/r1_prof_wgs30x $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
tail -c 600 gpurun_out/r1_bench_default.json; tail -c 400 gpurun_out/r1_bench_reference.json
