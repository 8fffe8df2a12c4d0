# Round-1 profiling recipe (run under gpurun): launch list + full capture of the top kernels.
# ncu only runs after the identical command exited 0 without it.
set -x
CMD="python bench.py --workload chr21 --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e"
$CMD > gpurun_out/chr21_plain.json 2> gpurun_out/chr21_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_chr21.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_walk|k_depth_tiles|k_span_agg' -s 9 -c 3 -o gpurun_out/prof_r1b $CMD > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
