# narrow depth fetch: its parity test, the whole GPU suite, then the default bench (e2e with the narrow fetch and,
# beside it, the plain 32-bit DMA)
set -x
timeout 300 python -m pytest tests -m gpu -x -q -k "narrow_depth_fetch" 2>&1 | tail -15
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_fetch.json 2> gpurun_out/b_fetch.err
tail -3 gpurun_out/b_fetch.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b_fetch.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e'])
PY
