#!/bin/bash
# One optimisation iteration on the GPU box: parity tests, 256-pair timing, then an ncu capture of the kernel.
#   tools/gpu_iter.sh <profile-name> [kernel regex]
set -o pipefail
name=${1:-prof}; rx=${2:-warp_photo_stream}
python -m pytest tests/test_warp_photo_gpu.py -q -m gpu 2>&1 | grep -E "^FAILED|passed|failed" | cut -c1-150
python tools/time_vg.py 256 5 2>&1 | grep -E "pairs|rel diffs"
python tools/profile_step.py 32 3 vg > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -f -o gpurun_out/$name python tools/profile_step.py 32 3 vg > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/ncu.log
