set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or adversarial or prepass or refuses or multi_contig or pileups or empty or sharding or long_reads or streamed or pipelined" 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-full-map > gpurun_out/r2_b7.json 2> gpurun_out/r2_b7.err; tail -c 1500 gpurun_out/r2_b7.json; tail -5 gpurun_out/r2_b7.err
CSV_CLAIM_REFLEN=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/r2_b7_noclaim.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/r2_b7_noclaim.json'));print('noclaim', round(d['ms_per_step'],4), {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})"
DIAG_REPEAT=3 DIAG_PINNED=1 timeout 600 python scripts/diag_strong.py 2 wgs30x > gpurun_out/r2_diag_strong2.txt 2>&1; grep -v "^contig" gpurun_out/r2_diag_strong2.txt | tail; grep -c DIFFERS gpurun_out/r2_diag_strong2.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
