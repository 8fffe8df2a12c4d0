set -x
mkdir -p gpurun_out
timeout 600 python scripts/stress_shard.py 2 1 100 > gpurun_out/r2_stress_2_1.txt 2>&1; tail -60 gpurun_out/r2_stress_2_1.txt
