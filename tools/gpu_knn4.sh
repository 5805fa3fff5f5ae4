#!/bin/bash
python tools/time_knn.py 307200 2000000 0.15 x
ncu --set full --clock-control none --import-source on -k regex:kg_query_far -s 1 -c 1 -f -o gpurun_out/prof_r2_knnfar python tools/time_knn.py 307200 2000000 0.15 x > gpurun_out/ncu_knnfar.log 2>&1
tail -2 gpurun_out/ncu_knnfar.log
