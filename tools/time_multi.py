"""S = 2 source frames per target (config C5 scale 0, 1080x1920): one multi-source launch vs one launch per source frame."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "end-to-end-self-supervised-slam_b200")); sys.path.insert(0, ROOT)
import e2e_slam_b200 as e2e
from e2e_slam_b200 import ops
from e2e_slam_b200.synthetic import make_pairs
B, H, W, S = int(sys.argv[1]) if len(sys.argv) > 1 else 16, 1080, 1920, 2
dev = torch.device("cuda:0")
ds = [make_pairs(B, H, W, "icl", seed=s, device=dev) for s in range(S)]
tgt = ds[0]["colors"][:, 1].permute(0, 3, 1, 2)
srcs = torch.stack([d["colors"][:, 0] for d in ds], 1)
Ts = torch.stack([d["T"] for d in ds], 1)
depth, inv_K, K = ds[0]["depth"], ds[0]["inv_K"], ds[0]["K"]


def multi():
    d = depth.detach().requires_grad_(True)
    s = srcs.detach().requires_grad_(True)
    ops.warp_photometric_loss_multi(d, inv_K, K, Ts, s.permute(0, 1, 4, 2, 3), tgt).backward()


def per_frame():
    d = depth.detach().requires_grad_(True)
    s = srcs.detach().requires_grad_(True)
    (sum(e2e.warp_photometric_loss(d, inv_K, K, Ts[:, i], s[:, i].permute(0, 3, 1, 2), tgt) for i in range(S)) / S).backward()


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


tm, tp = t(multi), t(per_frame)
npx = B * S * H * W
print(f"B={B} S={S} {H}x{W}: multi-source launch {tm:.3f} ms ({npx / tm / 1e6:.2f} Gpx/s, {108 * B * H * W / tm / 1e6:.0f} GB/s at 108 B/px per target px), "
      f"per-frame launches {tp:.3f} ms ({npx / tp / 1e6:.2f} Gpx/s)")
