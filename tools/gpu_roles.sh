#!/bin/bash
# role-split streaming kernel vs the classic one: parity tests (both kernels), 256-pair / 32-pair / 1080p timing
python -m pytest tests/test_round2_gpu.py -q -m gpu -k role_split 2>&1 | grep -E "^FAILED|passed|failed|Error|assert" | cut -c1-200
E2E_ROLES=1 python -m pytest tests/test_warp_photo_gpu.py tests/test_fullsize_gpu.py -q -m gpu -x 2>&1 | grep -E "^FAILED|passed|failed|Error" | cut -c1-200
for r in 1 0; do
echo "== E2E_ROLES=$r"
E2E_ROLES=$r python tools/time_vg.py 256 10 2>&1 | grep -E "pairs|rel diffs"
E2E_ROLES=$r python tools/time_vg.py 32 10 2>&1 | grep -E "pairs"
E2E_ROLES=$r python tools/time_vg.py 32 5 1080 1920 2>&1 | grep -E "pairs"
done
