#!/bin/bash
# scaling run as the driver does it: N GPUs, strong (default) with the weak leg alongside
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err
echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n${N}.json').read().strip().splitlines()[-1])
    print("N=${N} strong: value %.2f Gpx/s ms/step %.3f kernel_ms %.3f | other %s | e2e %.2f Gpx/s" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["ms_per_launch"], json.dumps(d.get("weak_scaling"))[:260], d["e2e"]["value"]/1e9))
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/bench_n${N}.err').read()[-2000:])
PY
