"""Small fixed workload for ncu: a few forward+backward passes of the fused kernels (and, later, fusion)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "end-to-end-self-supervised-slam_b200")); sys.path.insert(0, ROOT)
from e2e_slam_b200 import ops
from e2e_slam_b200.synthetic import make_pairs
P = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
H, W = 480, 640
dev = torch.device("cuda:0")
d = make_pairs(P, H, W, "icl", seed=0, device=dev)
src, tgt = d["colors"][:, 0].permute(0, 3, 1, 2), d["colors"][:, 1].permute(0, 3, 1, 2)
plan = ops.WarpPhotoPlan(P, H, W, dev)
mode = sys.argv[3] if len(sys.argv) > 3 else "vg"
for _ in range(iters):
    if mode == "vg":
        plan.value_and_grad(d["depth"], d["inv_K"], d["K"], d["T"], src, tgt)
    else:
        plan.forward(d["depth"], d["inv_K"], d["K"], d["T"], src, tgt)
        plan.backward(d["depth"], d["inv_K"], d["K"], d["T"], src, tgt)
torch.cuda.synchronize()
print("ok", float(plan.loss))
