mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
run() { name=$1; shift; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e "$@" > gpurun_out/r2_b39_$name.json 2>gpurun_out/r2_b39_$name.err; python -c "
import json;d=json.load(open('gpurun_out/r2_b39_$name.json'));print('$name', round(d['ms_per_step'],4), round(d['roofline']['path']['frac'],4), {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['gpu_launches'], d['retimed'])"; tail -2 gpurun_out/r2_b39_$name.err; }
run wgs
run svrich --workload svrich_wgs
CSV_SIG_ORDER=radix timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or adversarial or multi_contig or pileup or one_shot" 2>&1 | tail -3
