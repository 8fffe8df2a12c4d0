# GPU side of the round's evidence: default bench line, reference arm, launch list, ncu --set full of the hot kernels
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e"
$CMD > gpurun_out/r2_plain.json 2> gpurun_out/r2_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_wgs30x.csv $CMD > gpurun_out/r2_ncu_ll.log 2>&1
$CMD > /dev/null 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_depth_tiles16|k_walk|k_span_carry|k_pmax_chained|k_prep_dense' -s 10 -c 6 -f -o gpurun_out/r2_prof_wgs30x $CMD > gpurun_out/r2_ncu_full.log 2>&1
tail -3 gpurun_out/r2_ncu_full.log
ls -la gpurun_out/*.ncu-rep
