"""Min-reprojection objective (2 source frames, 64 targets): speculative forward + conditional backward against the
forward kernel + gradient-map sweep that expect_uniform=False selects."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "end-to-end-self-supervised-slam_b200")); sys.path.insert(0, ROOT)
from e2e_slam_b200 import losses, ops
from e2e_slam_b200.synthetic import make_pairs
B, H, W = 64, 480, 640
dev = torch.device("cuda:0")
ds = [{k: torch.cat([make_pairs(32, H, W, "icl", seed=10 * s + c, device=dev)[k] for c in range(B // 32)]) for k in ("depth", "inv_K", "K", "T", "colors")} for s in range(2)]
tgt = ds[0]["colors"][:, 1].permute(0, 3, 1, 2)
for uniform in (True, False):
    def step():
        d = ds[0]["depth"].detach().requires_grad_(True)
        maps = [ops.warp_photometric(d, ds[0]["inv_K"], ds[0]["K"], ds[s]["T"], ds[s]["colors"][:, 0].permute(0, 3, 1, 2), tgt, expect_uniform=uniform) for s in range(2)]
        loss = losses.photometric_objective(maps, min_reprojection=True)
        loss.backward()
        return loss, d.grad
    for _ in range(3): l, g = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): l, g = step()
    e1.record(); torch.cuda.synchronize()
    print(f"expect_uniform={uniform}: {e0.elapsed_time(e1) / 10:.3f} ms per step (64 targets x 2 sources, min-reprojection), loss {float(l):.8f} |grad| {float(g.abs().sum()):.6f}")
