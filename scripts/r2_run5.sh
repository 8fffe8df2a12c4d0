set -x
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/r2_b5_$name.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/r2_b5_$name.json'));print('$name', round(d['ms_per_step'],4), {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})"; }
run base X=1
run chunks2 CSV_CHUNKS=2
run chunks3 CSV_CHUNKS=3
run chunks4 CSV_CHUNKS=4
run chunks6 CSV_CHUNKS=6
run chunks8 CSV_CHUNKS=8
run chunks4_p0 CSV_CHUNKS=4 CSV_SIDE_PRIO=0
run walk64 CSV_WALK_GRID=64
run walk256 CSV_WALK_GRID=256
