set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or adversarial or prepass or refuses or multi_contig or pileups or empty or sharding or long_reads or streamed_shards or pipelined or mirror or one_shot or compiled_reference" 2>&1 | tail -8
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/r2_b14_$name.json 2>gpurun_out/r2_b14_$name.err; python -c "
import json;d=json.load(open('gpurun_out/r2_b14_$name.json'));print('$name', round(d['ms_per_step'],4), round(d['roofline']['path']['frac'],4), {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})"; }
run base X=1
run minb5 CSV_WALK_MINB=5
run base2 X=1
