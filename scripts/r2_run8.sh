set -x
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/r2_b8_$name.json 2>gpurun_out/r2_b8_$name.err; python -c "
import json;d=json.load(open('gpurun_out/r2_b8_$name.json'));print('$name', round(d['ms_per_step'],4), {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})"; }
run claim X=1
run noclaim CSV_CLAIM_REFLEN=0
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "prepass or pileup_at_one or ont60x or long_cigar or compiled_reference" 2>&1 | tail -8
bash scripts/r2_run_strong.sh 2
