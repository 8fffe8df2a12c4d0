# usage: bash scripts/gpu_sweep.sh "ENV1=a ENV2=b" "ENV1=c" ...   -- one short bench per environment setting
mkdir -p gpurun_out
i=0
for e in "$@"; do
  i=$((i+1))
  env $e python bench.py --steps 8 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/sw_$i.json 2> gpurun_out/sw_$i.err
  python - "$e" gpurun_out/sw_$i.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    st=d['roofline']['stage_ms_per_step']
    print('%-40s %.3f ms  '%(sys.argv[1], d['ms_per_step']), ' '.join('%s=%.3f'%(k,v) for k,v in st.items()))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
done
