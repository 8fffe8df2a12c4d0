mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/r2_b31_$name.json 2>gpurun_out/r2_b31_$name.err; python -c "
import json;d=json.load(open('gpurun_out/r2_b31_$name.json'));print('$name', round(d['ms_per_step'],4), round(d['roofline']['path']['frac'],4), {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['gpu_launches'])"; tail -2 gpurun_out/r2_b31_$name.err; }
run fusedprep X=1
run nofuse CSV_PREP_FUSED=0
STRESS_TIME=1 timeout 300 python scripts/stress_shard.py 8 3 2>&1 | tail -1
