#!/bin/bash
# N = 2: strong (default) + weak scaling legs, NCCL CTA caps
mkdir -p gpurun_out
for ctas in 0 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --nccl-max-ctas $ctas > gpurun_out/bench_n2_c$ctas.json 2> gpurun_out/bench_n2_c$ctas.err
  echo "ctas=$ctas rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n2_c$ctas.json').read().strip().splitlines()[-1])
    print(" strong: value %.2f Gpx/s ms/step %.3f kernel_ms %.3f | other %s | e2e %.2f Gpx/s" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["ms_per_launch"], json.dumps(d.get("weak_scaling"))[:200], d["e2e"]["value"]/1e9))
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/bench_n2_c$ctas.err').read()[-1500:])
PY
done
