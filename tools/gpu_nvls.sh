#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/nvls_check.py 2>&1 | grep -v "^W\|^\[W\|warn" | tail -8
bash tools/gpu_n8.sh $N
