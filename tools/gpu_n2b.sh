#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_round2_gpu.py -x -q -m gpu 2>&1 | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_v.json 2> gpurun_out/bench_n2_v.err
echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n2_v.json').read().strip().splitlines()[-1])
    print(" strong: value %.2f Gpx/s ms/step %.3f kernel_ms %.3f | other %s | e2e %.2f Gpx/s" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["ms_per_launch"], json.dumps(d.get("weak_scaling"))[:220], d["e2e"]["value"]/1e9))
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/bench_n2_v.err').read()[-1500:])
PY
