# round-2 evidence on one GPU: tests, profile pass (launch list + ncu --set full), default bench line, reference arm, other workloads
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash scripts/r2_profile.sh
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -c 600 gpurun_out/r2_bench_default.json; tail -3 gpurun_out/r2_bench_default.err
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; tail -c 800 gpurun_out/r2_bench_reference.json
timeout 600 python bench.py --workload svrich_wgs --no-cpu-baseline --no-full-map > gpurun_out/r2_bench_svrich_wgs.json 2> gpurun_out/r2_bench_svrich_wgs.err; tail -c 400 gpurun_out/r2_bench_svrich_wgs.json
timeout 600 python bench.py --workload ont60x_chr20 --no-cpu-baseline --no-full-map > gpurun_out/r2_bench_ont60x_chr20.json 2> gpurun_out/r2_bench_ont60x_chr20.err; tail -c 400 gpurun_out/r2_bench_ont60x_chr20.json
