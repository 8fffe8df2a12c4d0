set -x
mkdir -p gpurun_out
timeout 600 python scripts/stress_shard.py 2 1 200 > gpurun_out/r2_stress_2_1.txt 2>&1; tail -8 gpurun_out/r2_stress_2_1.txt
timeout 600 python scripts/stress_shard.py 1 0 60 > gpurun_out/r2_stress_1_0.txt 2>&1; tail -8 gpurun_out/r2_stress_1_0.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
