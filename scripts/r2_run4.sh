set -x
mkdir -p gpurun_out
timeout 600 python scripts/diag_strong.py 2 wgs30x > gpurun_out/r2_diag_strong.txt 2>&1; tail -30 gpurun_out/r2_diag_strong.txt
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/r2_b4_$name.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/r2_b4_$name.json'));print('$name', round(d['ms_per_step'],4), {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})"; }
run base X=1
run ctas16 CSV_SIDE_CTAS=16
run ctas32 CSV_SIDE_CTAS=32
run ctas64 CSV_SIDE_CTAS=64
run ctas148 CSV_SIDE_CTAS=148
run ctas32_p0 CSV_SIDE_CTAS=32 CSV_SIDE_PRIO=0
run ctas32_g192 CSV_SIDE_CTAS=32 CSV_TILE_GRID=192
run ctas32_g48 CSV_SIDE_CTAS=32 CSV_TILE_GRID=48
run g192 CSV_TILE_GRID=192
CONTEXTSV_B200_STATS=1 timeout 900 python -m pytest tests/test_dropin_cli.py -m gpu -x -q 2>&1 | tail -30
