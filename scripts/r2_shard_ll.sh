mkdir -p gpurun_out
STRESS_TIME=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 60 --csv --log-file gpurun_out/r2_shard8_ll.csv python scripts/stress_shard.py 8 3 > gpurun_out/r2_shard8_ll.log 2>&1
python - <<'P'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2_shard8_ll.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
col={h:i for i,h in enumerate(rows[hi])}
agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<len(col): continue
    n=r[col['Kernel Name']].split('(')[0][:46]; v=float(r[col['Metric Value']].replace(',',''))/1000.0
    agg.setdefault(n,[]).append(v)
for k,v in agg.items(): print('%-48s n=%2d mean %7.2f us'%(k,len(v),sum(v)/len(v)))
P
