#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_round2_gpu.py -q -m gpu -k "single_launch" 2>&1 | grep -E "^E  |passed|failed" | head -12
