#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests/test_round2_gpu.py tests/test_fullsize_gpu.py -q -m gpu 2>&1 | grep -E "passed|failed|^FAILED|^E  |Error" | head -40
cat gpurun_out/parity_measured.jsonl
