#!/bin/bash
# classic streaming kernel with TMA-staged rows: parity, timing, ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_warp_photo_gpu.py tests/test_configs_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== TMA classic"; timeout 300 python tools/time_vg.py 256 10
timeout 200 python tools/profile_step.py 32 3 vg > gpurun_out/plain_x.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp_photo_stream_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2c python tools/profile_step.py 32 3 vg > gpurun_out/ncu_x.log 2>&1
tail -2 gpurun_out/ncu_x.log
