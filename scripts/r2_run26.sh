set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 1500 python scripts/bench_cli.py --mbp 46.7 --threads 16 --reps 2 > gpurun_out/r2_cli_chr21.json 2> gpurun_out/r2_cli_chr21_stats.txt; tail -c 1500 gpurun_out/r2_cli_chr21.json; tail -25 gpurun_out/r2_cli_chr21_stats.txt
