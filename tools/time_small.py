"""Latency of the single-sweep call for small batches (configs C1 / C2 are single pairs)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "end-to-end-self-supervised-slam_b200")); sys.path.insert(0, ROOT)
from e2e_slam_b200 import ops
from e2e_slam_b200.synthetic import make_pairs
dev = torch.device("cuda:0")
H, W = 480, 640
out = []
for P in (1, 2, 4, 8):
    d = make_pairs(P, H, W, "icl", seed=1, device=dev)
    src, tgt = d["colors"][:, 0].permute(0, 3, 1, 2), d["colors"][:, 1].permute(0, 3, 1, 2)
    plan = ops.WarpPhotoPlan(P, H, W, dev)
    a = (d["depth"], d["inv_K"], d["K"], d["T"], src, tgt)
    for _ in range(5):
        plan.value_and_grad(*a)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        plan.value_and_grad(*a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    out.append(f"P={P}: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us")
print(os.environ.get("E2E_S_MINSEG", "48"), " | ".join(out))
