#!/bin/bash
# full GPU test suite + default bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 900 python bench.py > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r2_a.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_a.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","scaling","e2e","roofline","cpu_baseline","torch_cuda_baseline","single_pair","point_supervision","gpu_launches","clocks"):
    print(k, json.dumps(d.get(k))[:400])
print("fusion", json.dumps(d.get("fusion"))[:300]); print("c2", json.dumps(d.get("c2_refinement_step"))[:300])
PY
