set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/b_t1.json 2> gpurun_out/b_t1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b_t1.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['stage_ms_per_step'])
PY
