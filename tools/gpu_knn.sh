#!/bin/bash
# kNN iteration on the GPU box: grid-vs-brute parity tests, far / near query timings, the C2 step.
python -m pytest tests/test_fusion_gpu.py -q -m gpu -k "knn or point_supervision or grid" 2>&1 | grep -E "^FAILED|passed|failed|Error" | cut -c1-200
python tools/time_knn.py 307200 2000000 0.0
python tools/time_knn.py 307200 2000000 0.02 x
python tools/time_knn.py 307200 2000000 0.09 x
python tools/time_knn.py 307200 2000000 0.2 x
python tools/time_knn.py 19200 75000 0.05
python tools/run_c2.py 2>&1 | tail -3
