set -x
CMD="python bench.py --workload chr1_5 --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e"
$CMD > gpurun_out/p_plain.json 2> gpurun_out/p_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-k_depth_tiles}" -s ${SKIP:-3} -c ${COUNT:-1} -f -o gpurun_out/prof_it $CMD > gpurun_out/ncu_it.log 2>&1
tail -2 gpurun_out/ncu_it.log
