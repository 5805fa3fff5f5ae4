#!/bin/bash
# launch list of the bench command (share of the step per kernel) + traffic of the dominant kernel; each only after the plain run exited 0
python bench.py --steps 2 --warmup 3 --skip-cpu --skip-torch-cuda > gpurun_out/b_plain.log 2>&1 || { echo plain failed; tail -3 gpurun_out/b_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-torch-cuda > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log | cut -c1-200
python - <<'P'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_r2.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[ii],{'k':r[ki]})[r[mi]]=float(r[vi].replace(',',''))
agg=collections.OrderedDict()
for i,v in d.items():
    a=agg.setdefault(v['k'][:90],[0,0.0,0.0,0.0]); a[0]+=1; a[1]+=v.get('gpu__time_duration.sum',0)/1e3; a[2]+=v.get('dram__bytes_read.sum',0); a[3]+=v.get('dram__bytes_write.sum',0)
print(f"{'kernel':92s} {'n':>4s} {'us total':>10s} {'us/launch':>10s}")
for k,a in sorted(agg.items(), key=lambda x:-x[1][1])[:40]:
    print(f"{k:92s} {a[0]:4d} {a[1]:10.1f} {a[1]/a[0]:10.1f}")
P
