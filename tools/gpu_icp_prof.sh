#!/bin/bash
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/icp_launches.csv python tools/time_icp.py > gpurun_out/icp_ncu.log 2>&1
python - <<'P'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/icp_launches.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=collections.OrderedDict()
seq=[]
for r in rows[1:]:
    k=r[ki].split('(')[0].replace('void ','').replace('e2e::','')[:40]; v=float(r[vi].replace(',',''))/1e3
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v; seq.append((k,v))
for k,a in sorted(agg.items(), key=lambda x:-x[1][1])[:16]:
    print(f"{k:42s} n={a[0]:4d} total {a[1]:9.1f} us  avg {a[1]/a[0]:7.1f}")
# one gradicp iteration: find first 'icp_solve_kernel' occurrences pattern
idx=[i for i,(k,v) in enumerate(seq) if k.startswith('icp_solve')]
if len(idx)>30:
    a,b=idx[25],idx[27]
    print("one GradICP iteration (between solve launches):")
    for k,v in seq[a+1:b+1]: print(f"   {k:42s} {v:7.1f} us")
P
