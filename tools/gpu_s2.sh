#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_round2_gpu.py -q -m gpu -k "inside_its_buffers or batched" 2>&1 | tail -3
timeout 600 python tools/time_fusion.py 2>&1 | tail -2 | cut -c1-1500
timeout 300 python tools/profile_fusion.py 60 4 > gpurun_out/plain_fb.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fusion_sequence_batch -s 1 -c 1 -f -o gpurun_out/prof_r2_fusion_b4 python tools/profile_fusion.py 60 4 > gpurun_out/ncu_fb.log 2>&1
tail -2 gpurun_out/ncu_fb.log
