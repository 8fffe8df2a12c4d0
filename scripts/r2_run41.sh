mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "more_signatures or golden or one_shot or pileup" 2>&1 | tail -4
CSV_STREAM_CONTIGS=1 timeout 900 python bench.py --workload ont60x_wgs --steps 3 > gpurun_out/r2_ont_wgs_try.json 2> gpurun_out/r2_ont_wgs_try.err; tail -c 1500 gpurun_out/r2_ont_wgs_try.json; tail -5 gpurun_out/r2_ont_wgs_try.err
