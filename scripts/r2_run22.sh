set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or adversarial or prepass or refuses or multi_contig or pileups or empty or sharding or long_reads or streamed_shards or pipelined or mirror or one_shot or compiled_reference or narrow" 2>&1 | tail -8
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/r2_b22_$name.json 2>gpurun_out/r2_b22_$name.err; python -c "
import json;d=json.load(open('gpurun_out/r2_b22_$name.json'));print('$name', round(d['ms_per_step'],4), round(d['roofline']['path']['frac'],4), {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})"; tail -2 gpurun_out/r2_b22_$name.err; }
run g4 X=1
run g8 CSV_TILE_GRID=8
run g16 CSV_TILE_GRID=16
run g96 CSV_TILE_GRID=96
timeout 300 python scripts/stress_shard.py 1 0 60 2>&1 | tail -3
timeout 300 python scripts/stress_shard.py 8 3 100 2>&1 | tail -3
