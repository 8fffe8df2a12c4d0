# Full ncu capture of the top kernels on a workload large enough to be in steady state (chr1-5, 1.06 Gb).
set -x
CMD="python bench.py --workload chr1_5 --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e"
$CMD > gpurun_out/chr1_5_plain.json 2> gpurun_out/chr1_5_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:'k_walk|k_depth_tiles|k_span_agg' -s 9 -c 3 -o gpurun_out/prof_big $CMD > gpurun_out/ncu_big.log 2>&1
tail -3 gpurun_out/ncu_big.log
