mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/r2_b32_$name.json 2>gpurun_out/r2_b32_$name.err; python -c "
import json;d=json.load(open('gpurun_out/r2_b32_$name.json'));print('$name', round(d['ms_per_step'],4), round(d['roofline']['path']['frac'],4), {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['gpu_launches'])"; tail -2 gpurun_out/r2_b32_$name.err; }
run base X=1
run wminb5 CSV_WALK_MINB=5
run tg64 CSV_TILE_GRID=64
run tg128 CSV_TILE_GRID=128
run tg192 CSV_TILE_GRID=192
run wg256 CSV_WALK_GRID=256
run base2 X=1
