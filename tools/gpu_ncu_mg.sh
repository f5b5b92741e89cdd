#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --loop-mode host --precond mgline > gpurun_out/plain_mg.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_mg.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --loop-mode host --precond mgline > gpurun_out/ncu_mg.log 2>&1
echo rc=$?
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_mg.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; data=rows[hi+1:]
kn,mv,mn,mu,gs=(hdr.index(x) for x in ('Kernel Name','Metric Value','Metric Name','Metric Unit','Grid Size'))
agg=collections.OrderedDict()
for r in data:
    if len(r)<=mv or r[mn]!='gpu__time_duration.sum': continue
    v=float(r[mv].replace(',',''))*{'ns':1e-3,'us':1.0,'ms':1e3}.get(r[mu],1.0)
    name=r[kn].split('(')[0].replace('void ','').replace('<unnamed>::','')[:28]+' grid='+r[gs]
    c=agg.setdefault(name,[0,0.0]); c[0]+=1; c[1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:40]:
    print(f"{k:50s} n={v[0]:5d} avg={v[1]/v[0]:8.2f} us share={100*v[1]/tot:5.1f}%")
PY
