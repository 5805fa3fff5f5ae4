"""ICP / GradICP odometry timing at the reference's sizes (live frame thinned 4x4 = 19 200 points vs ~75 000 active map points)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "end-to-end-self-supervised-slam_b200")); sys.path.insert(0, ROOT)
from e2e_slam_b200 import odometry
g = torch.Generator(device="cuda").manual_seed(1)
M, N = 75000, 19200
uv = torch.rand(M, 2, generator=g, device="cuda") * 4 - 2
z = 2.5 + 0.3 * torch.sin(2 * uv[:, 0]) * torch.cos(uv[:, 1])
tgt = torch.stack([uv[:, 0], uv[:, 1], z], 1)
n = torch.stack([-0.6 * torch.cos(2 * uv[:, 0]) * torch.cos(uv[:, 1]), 0.3 * torch.sin(2 * uv[:, 0]) * torch.sin(uv[:, 1]), torch.ones(M, device="cuda")], 1)
n = n / n.norm(dim=1, keepdim=True)
src = tgt[torch.randperm(M, device="cuda", generator=g)[:N]] + torch.tensor([0.01, -0.015, 0.02], device="cuda")
eye = torch.eye(4, device="cuda")
for name, fn in (("icp", lambda: odometry.point_to_plane_ICP(src[None], tgt[None], n[None], eye, 20)),
                 ("gradicp", lambda: odometry.point_to_plane_gradICP(src[None], tgt[None], n[None], eye, 20, nu=0.05))):
    with torch.no_grad():
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            T, _ = fn()
        torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms per 20-iteration alignment, t = {T[:3, 3].tolist()}")

# with autograd: the library's loop + reverse sweep against the iteration written with torch ops
for name, fn in (("gradicp library fwd+bwd", odometry.point_to_plane_gradICP), ("gradicp torch-op fwd+bwd", odometry.point_to_plane_gradICP_torch),
                 ("icp library fwd+bwd", odometry.point_to_plane_ICP), ("icp torch-op fwd+bwd", odometry.point_to_plane_ICP_torch)):
    kw = dict(nu=0.05) if "gradicp" in name else {}
    def run():
        s = src[None].clone().requires_grad_(True)
        T, _ = fn(s, tgt[None], n[None], eye, 20, **kw)
        T[:3].sum().backward()
        return s.grad
    run(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        g = run()
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms, |grad| = {float(g.abs().sum()):.4f}")
