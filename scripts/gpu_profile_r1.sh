set -x
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
CMD="python bench.py --workload chr21 --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e"
$CMD > gpurun_out/chr21_plain.json 2> gpurun_out/chr21_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_chr21.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_walk|k_depth_tiles' -s 6 -c 4 -o gpurun_out/prof_r1a $CMD > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_wgs2.json 2> gpurun_out/bench_wgs2.err; tail -c 1500 gpurun_out/bench_wgs2.json
