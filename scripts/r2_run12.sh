set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -c 4000 gpurun_out/r2_bench_default.json; tail -3 gpurun_out/r2_bench_default.err
bash scripts/r2_profile.sh
