set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b1.json 2> gpurun_out/r2_b1.err; tail -c 3000 gpurun_out/r2_b1.json
CSV_REC_PREPASS=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/r2_b1_old.json 2> gpurun_out/r2_b1_old.err; tail -c 1500 gpurun_out/r2_b1_old.json
