set -x
CSV_BENCH_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --e2e-steps 2 --no-cpu-baseline > gpurun_out/b_fetch.json 2> gpurun_out/b_fetch.err
grep trace gpurun_out/b_fetch.err | head -24
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b_fetch.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], json.dumps(d['e2e']))
PY
