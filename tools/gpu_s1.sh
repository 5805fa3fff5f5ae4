#!/bin/bash
# memcheck over the reduced case set, then the one failing test again, then the fusion bench with the batched sweep
mkdir -p gpurun_out
bash tools/sanitize.sh memcheck
timeout 600 python -m pytest tests/test_round2_gpu.py -q -m gpu -k "patched or graphed" 2>&1 | tail -3
timeout 600 python tools/time_fusion.py 2>&1 | tail -3 | cut -c1-1500
