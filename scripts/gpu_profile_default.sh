# Evidence for the default bench command (config 2): launch list (device time per launch) and one
# --set full capture of the dominant kernel.  ncu only runs after the identical command exited 0 without it.
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e"
$CMD > gpurun_out/wgs_plain.json 2> gpurun_out/wgs_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_wgs.csv $CMD > gpurun_out/ncu_wgs1.log 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_depth_tiles|k_walk|k_span_agg' -s 9 -c 3 -o gpurun_out/prof_wgs $CMD > gpurun_out/ncu_wgs2.log 2>&1
tail -2 gpurun_out/ncu_wgs2.log
