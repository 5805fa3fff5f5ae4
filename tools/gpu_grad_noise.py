"""Developer diagnostic (GPU): three-way gradient error table  ours / reference fp32 / reference fp64."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "end-to-end-self-supervised-slam_b200")); sys.path.insert(0, ROOT)
import e2e_slam_b200 as e2e
from e2e_slam_b200.synthetic import make_pairs
from oracle import torch_oracle
def l2(a, b): return float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))
def mx(a, b): return float(np.abs(a.astype(np.float64) - b).max() / np.abs(b).max())
for (B, H, W, kind, pad, mask, rot, trans) in [(1, 16, 64, "icl", "border", False, 1.0, 0.02), (1, 97, 131, "icl", "border", True, 2.0, 0.05),
                                                (1, 480, 640, "icl", "border", True, 2.0, 0.05), (1, 480, 640, "tum", "border", True, 5.0, 0.15),
                                                (2, 64, 64, "tum", "zeros", True, 5.0, 0.30)]:
    d = make_pairs(B, H, W, kind, seed=H * 1000 + W, rot_deg=rot, trans=trans)
    r32 = torch_oracle.fwd_bwd(d["depth"], d["inv_K"], d["K"], d["T"], d["colors"][:, 0], d["colors"][:, 1], pad, mask)
    r64 = torch_oracle.fwd_bwd(d["depth"], d["inv_K"], d["K"], d["T"], d["colors"][:, 0], d["colors"][:, 1], pad, mask, dtype=torch.float64)
    dev = {k: v.cuda() for k, v in d.items()}
    depth = dev["depth"].clone().requires_grad_(True); colors = dev["colors"].clone().requires_grad_(True); T = dev["T"].clone().requires_grad_(True)
    lm = e2e.warp_photometric(depth, dev["inv_K"], dev["K"], T, colors[:, 0].permute(0, 3, 1, 2), colors[:, 1].permute(0, 3, 1, 2), pad, mask)
    lm.mean().backward()
    print(f"--- {B}x{H}x{W} {kind} {pad} rot {rot} trans {trans}")
    for k, o in (("g_depth", depth.grad), ("g_src", colors.grad[:, 0]), ("g_T", T.grad)):
        o = o.cpu().numpy(); a32 = r32[k].numpy(); a64 = r64[k].numpy()
        print(f"  {k:8s} ours-ref32 L2 {l2(o, a32.astype(np.float64)):.2e} max {mx(o, a32.astype(np.float64)):.2e} | ours-ref64 L2 {l2(o, a64):.2e} max {mx(o, a64):.2e} | ref32-ref64 L2 {l2(a32, a64):.2e} max {mx(a32, a64):.2e}")
