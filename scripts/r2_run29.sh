mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/r2_b29_$name.json 2>gpurun_out/r2_b29_$name.err; python -c "
import json;d=json.load(open('gpurun_out/r2_b29_$name.json'));print('$name', round(d['ms_per_step'],4), round(d['roofline']['path']['frac'],4), {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['gpu_launches'])"; tail -2 gpurun_out/r2_b29_$name.err; }
run base X=1
run ctas74 CSV_SIDE_CTAS=74
run ctas148 CSV_SIDE_CTAS=148
run ctas296 CSV_SIDE_CTAS=296
run ctas592 CSV_SIDE_CTAS=592
run grid2 CSV_SIDE_GRID=2
for c in 0 74 148 296; do echo "CSV_SIDE_CTAS=$c"; CSV_SIDE_CTAS=$c STRESS_TIME=1 timeout 300 python scripts/stress_shard.py 8 3 2>&1 | tail -1; done
