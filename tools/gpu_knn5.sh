#!/bin/bash
for v in 7 8 10 12; do
echo "== M: ${v:-default 4}"
if [ -n "$v" ]; then export E2E_LIB_PATH=end-to-end-self-supervised-slam_b200/lib/variant_m$v.so; fi
python -m pytest tests/test_fusion_gpu.py -q -m gpu -k "grid" 2>&1 | grep -E "^FAILED|passed|failed|Error" | cut -c1-200
python tools/time_knn.py 307200 2000000 0.0 x
python tools/time_knn.py 307200 2000000 0.02 x
python tools/time_knn.py 307200 2000000 0.09 x
python tools/time_knn.py 307200 2000000 0.15 x
python tools/time_knn.py 19200 75000 0.05 x
python tools/run_c2.py 2>&1 | grep -E "ms_per_step_cuda" 
done
