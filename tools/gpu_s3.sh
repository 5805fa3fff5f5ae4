#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_warp_photo_gpu.py tests/test_configs_gpu.py -q -m gpu 2>&1 | tail -4
timeout 300 python tools/time_multi.py 16
timeout 300 python tools/time_vg.py 256 10 | tail -2
