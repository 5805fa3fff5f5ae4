#!/bin/bash
# one micro-optimisation iteration of the classic streaming kernel: parity, 256-pair time, executed warp instructions (32 pairs)
python -m pytest tests/test_warp_photo_gpu.py -q -m gpu -x 2>&1 | grep -E "^FAILED|passed|failed|Error" | cut -c1-200
python tools/time_vg.py 256 10 2>&1 | grep -E "pairs|rel diffs"
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:warp_photo_stream -s 1 -c 1 python tools/profile_step.py 32 3 vg 2>&1 | grep -E "inst_executed|time_duration"
