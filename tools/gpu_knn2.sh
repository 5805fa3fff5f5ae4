#!/bin/bash
python tools/run_c2.py 2>&1 | grep -E "ms_per_step|launches" 
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/knn_launches.csv python tools/time_knn.py 307200 2000000 0.09 x > gpurun_out/ncu_knn.log 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/knn_launches.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[1:]:
    print(r[ki][:60], r[vi])
P
