set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'k_walk|k_span_carry|k_chained_scan|k_prep_dense' -c 24 --csv --log-file gpurun_out/r2_walk_ll.csv $CMD > gpurun_out/r2_walk_ll.log 2>&1
python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/r2_walk_ll.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
col={h:i for i,h in enumerate(rows[hi])}
for r in rows[hi+1:]:
    if len(r)>=len(col): print(r[col['Kernel Name']][:50], r[col['Metric Name']], r[col['Metric Value']], r[col['Metric Unit']])
P
