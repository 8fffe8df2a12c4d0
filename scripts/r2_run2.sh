set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or adversarial or prepass or refuses or multi_contig or pileups or empty or sharding or long_reads" 2>&1 | tail -15
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b2.json 2> gpurun_out/r2_b2.err; tail -c 2500 gpurun_out/r2_b2.json; tail -5 gpurun_out/r2_b2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_wgs30x.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/r2_ncu.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
