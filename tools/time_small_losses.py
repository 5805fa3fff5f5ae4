"""Achieved bandwidth of the small losses (smoothness fwd+bwd, sparse-gt L1, regulariser) at 256 x 480 x 640."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "end-to-end-self-supervised-slam_b200")); sys.path.insert(0, ROOT)
from e2e_slam_b200 import losses
B, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 480, 640
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
disp = (torch.rand(B, 1, H, W, generator=g, device=dev) + 0.1)
img = torch.rand(B, H, W, 3, generator=g, device=dev).permute(0, 3, 1, 2)      # NCHW view of NHWC memory, like the scripts
gt = torch.rand(B, H, W, 1, generator=g, device=dev)
mask = (torch.rand(B, H, W, 1, generator=g, device=dev) < 0.012).float()
init = torch.rand(B, 1, H, W, generator=g, device=dev)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def smooth():
    d = disp.detach().requires_grad_(True)
    losses.smoothness_loss(d, img).backward()


def sparse():
    d = disp.detach().requires_grad_(True)
    losses.depth_gt_loss(d[:1], gt[:1], mask[:1]).backward() if B == 1 else None


def reg():
    d = disp.detach().requires_grad_(True)
    losses.depth_reguralizer(init, d, "l1").backward()


npx = B * H * W
for name, fn, bpp in (("smoothness fwd+bwd", smooth, 40), ("depth_reguralizer l1 fwd+bwd", reg, 16)):
    t = timeit(fn)
    print(f"{name}: {t:.3f} ms, {bpp * npx / t / 1e6:.0f} GB/s algorithmic ({bpp} B/px)")
