# ncu --set full of the dominant kernel alone (one launch) -> gpurun_out/r2_prof_tiles.ncu-rep; then the default bench line
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_depth_tiles16' -s 3 -c 1 -f -o gpurun_out/r2_prof_tiles $CMD > gpurun_out/r2_ncu_tiles.log 2>&1
tail -2 gpurun_out/r2_ncu_tiles.log
