"""Where does the streaming kernel's time go?  256 pairs with / without the grad_src scatter (need_src_grad=False passes grad_src = NULL)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "end-to-end-self-supervised-slam_b200")); sys.path.insert(0, ROOT)
from e2e_slam_b200 import ops
from e2e_slam_b200.synthetic import make_pairs
P = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H, W = 480, 640
dev = torch.device("cuda:0")
chunks = [make_pairs(min(32, P - s), H, W, "icl", seed=s, device=dev) for s in range(0, P, 32)]
d = {k: torch.cat([c[k] for c in chunks]) for k in chunks[0]}
src, tgt = d["colors"][:, 0].permute(0, 3, 1, 2), d["colors"][:, 1].permute(0, 3, 1, 2)
a = (d["depth"], d["inv_K"], d["K"], d["T"], src, tgt)
for need in (True, False):
    plan = ops.WarpPhotoPlan(P, H, W, dev, need_src_grad=need)
    for _ in range(3):
        plan.value_and_grad(*a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        plan.value_and_grad(*a)
    e1.record(); torch.cuda.synchronize()
    print(f"pairs={P} need_src_grad={need}: {e0.elapsed_time(e1) / 5:.3f} ms")
