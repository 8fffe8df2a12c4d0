set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_walk' -s 3 -c 1 -f -o gpurun_out/r2_prof_walk $CMD > gpurun_out/r2_walk_full.log 2>&1
tail -2 gpurun_out/r2_walk_full.log
