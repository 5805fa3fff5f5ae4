"""Synthetic ICL-NUIM / TUM-shaped key-frame pairs: there is no dataset on the
benchmark box, so bench.py, smoke() and the tests build inputs of the reference's shapes and layouts:
colours channels-last (B, L, H, W, 3) in [0,1] as gradslam's loaders yield them after `/255`
(train_depth.py:254), depth (B,1,H,W) in metres, intrinsics and relative pose as (B,4,4)."""
import math

import torch

ICL_K = (481.2, -480.0, 319.5, 239.5)   # fx, fy (negative in ICL-NUIM), cx, cy at 640x480
TUM_K = (525.0, 525.0, 319.5, 239.5)


def intrinsics(kind, B, H, W, device="cpu"):
    fx, fy, cx, cy = ICL_K if kind == "icl" else TUM_K
    K = torch.eye(4, dtype=torch.float32).repeat(B, 1, 1)
    K[:, 0, 0], K[:, 1, 1] = fx * W / 640.0, fy * H / 480.0
    K[:, 0, 2], K[:, 1, 2] = (cx + 0.5) * W / 640.0 - 0.5, (cy + 0.5) * H / 480.0 - 0.5
    return K.to(device)


def se3_exp(w, t):
    """Rodrigues: (B,3) axis-angle, (B,3) translation -> (B,4,4)."""
    B = w.shape[0]
    th = w.norm(dim=1, keepdim=True).clamp_min(1e-12)
    k = w / th
    Kx = torch.zeros(B, 3, 3, dtype=w.dtype)
    Kx[:, 0, 1], Kx[:, 0, 2] = -k[:, 2], k[:, 1]
    Kx[:, 1, 0], Kx[:, 1, 2] = k[:, 2], -k[:, 0]
    Kx[:, 2, 0], Kx[:, 2, 1] = -k[:, 1], k[:, 0]
    th = th.unsqueeze(-1)
    R = torch.eye(3, dtype=w.dtype) + torch.sin(th) * Kx + (1 - torch.cos(th)) * (Kx @ Kx)
    T = torch.eye(4, dtype=w.dtype).repeat(B, 1, 1)
    T[:, :3, :3], T[:, :3, 3] = R, t
    return T


def make_pairs(B, H, W, kind="icl", seed=0, device="cpu", rot_deg=2.0, trans=0.05, holes=0.0, frames=2):
    """Returns dict(depth, K, inv_K, T, colors) on `device`.  depth: tilted plane + ripples + noise in
    [~0.7, ~4] m; colours: sums of low-frequency sinusoids + noise; T: small random SE(3) motion
    (<= rot_deg degrees, <= trans metres).  `holes` = fraction of depth pixels set to 0 (TUM-style)."""
    device = torch.device(device)
    g = torch.Generator().manual_seed(seed)               # small parameters: CPU generator (same on any device)
    gd = torch.Generator(device=device).manual_seed(seed)   # per-pixel noise: generated where the data lives
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=device),
                            torch.arange(W, dtype=torch.float32, device=device), indexing="ij")
    a = (torch.rand(B, 2, 1, 1, generator=g) - 0.5).to(device)
    ph = (torch.rand(B, 1, 1, generator=g) * 6.28).to(device)
    depth = 2.0 + a[:, 0] * xs / W * 2 + a[:, 1] * ys / H * 2 + 0.3 * torch.sin(xs / (W / 12.0) + ph) * torch.cos(ys / (H / 9.0))
    depth = depth + 0.05 * torch.rand(B, H, W, generator=gd, device=device)
    if holes > 0:
        depth = depth * (torch.rand(B, H, W, generator=gd, device=device) >= holes)
    colors = torch.zeros(B, frames, H, W, 3, device=device)
    nw = 6
    kx = ((torch.rand(B, frames, 3, nw, generator=g) - 0.5) * 1.2 * 64.0 / W).to(device)
    ky = ((torch.rand(B, frames, 3, nw, generator=g) - 0.5) * 1.2 * 48.0 / H).to(device)
    p0 = (torch.rand(B, frames, 3, nw, generator=g) * 6.28).to(device)
    am = (torch.rand(B, frames, 3, nw, generator=g) * 0.8 + 0.2).to(device)
    for i in range(nw):
        colors += (am[..., i, None, None] * torch.sin(kx[..., i, None, None] * xs + ky[..., i, None, None] * ys
                                                      + p0[..., i, None, None])).permute(0, 1, 3, 4, 2)
    lo = colors.amin(dim=(2, 3), keepdim=True)
    hi = colors.amax(dim=(2, 3), keepdim=True)
    colors = 0.96 * (colors - lo) / (hi - lo + 1e-9) + 0.04 * torch.rand(B, frames, H, W, 3, generator=gd, device=device)
    w = torch.randn(B, 3, generator=g)
    w = w / w.norm(dim=1, keepdim=True) * math.radians(rot_deg) * (0.5 + 0.5 * torch.rand(B, 1, generator=g))
    t = torch.randn(B, 3, generator=g)
    t = t / t.norm(dim=1, keepdim=True) * trans * (0.5 + 0.5 * torch.rand(B, 1, generator=g))
    T = se3_exp(w, t)
    K = intrinsics(kind, B, H, W)
    inv_K = torch.pinverse(K)      # train_depth.py:460-461
    return dict(depth=depth.unsqueeze(1).contiguous().to(device), K=K.to(device), inv_K=inv_K.to(device),
                T=T.to(device), colors=colors.contiguous().to(device))


def room_sequence(L, H, W, device="cpu", seed=0):
    """Config C3: an analytic box room ray-cast to L depth + colour frames along a smooth trajectory with
    ground-truth poses (2-3 cm and ~0.6 degrees per frame).  Returns depth (L,H,W), rgb (L,H,W,3), K (4,4),
    poses (L,4,4) on `device`."""
    device = torch.device(device)
    g = torch.Generator().manual_seed(seed)
    fx = fy = 525.0 * W / 640.0
    cx, cy = W / 2 - 0.5, H / 2 - 0.5
    K = torch.eye(4)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2] = fx, fy, cx, cy
    half = torch.tensor([2.0, 1.4, 3.5], dtype=torch.float64, device=device)
    vv, uu = torch.meshgrid(torch.arange(H, dtype=torch.float64, device=device), torch.arange(W, dtype=torch.float64, device=device),
                            indexing="ij")
    rays = torch.stack([(uu - cx) / fx, (vv - cy) / fy, torch.ones_like(uu)], -1)
    depth = torch.empty(L, H, W, device=device)
    rgb = torch.empty(L, H, W, 3, device=device)
    poses = torch.empty(L, 4, 4)
    pos = torch.tensor([0.1, -0.1, 0.0], dtype=torch.float64)
    yaw = pitch = 0.0
    for s in range(L):
        cyw, syw, cp, sp = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch)
        Ry = torch.tensor([[cyw, 0, syw], [0, 1, 0], [-syw, 0, cyw]], dtype=torch.float64)
        Rx = torch.tensor([[1, 0, 0], [0, cp, -sp], [0, sp, cp]], dtype=torch.float64)
        R = Ry @ Rx
        dirs = rays @ R.t().to(device)
        p = pos.to(device)
        tt = torch.where(dirs > 0, (half - p) / dirs, (-half - p) / dirs)
        tt = torch.where(torch.isfinite(tt) & (tt > 0), tt, torch.full_like(tt, float("inf")))
        thit, wall = tt.min(-1)
        hit = p + dirs * thit.unsqueeze(-1)
        depth[s] = thit.float()
        tex = 0.5 + 0.25 * torch.sin(3.0 * hit[..., 0] + wall) + 0.25 * torch.cos(2.5 * hit[..., 1] + 1.3 * hit[..., 2])
        rgb[s] = torch.stack([tex, 0.8 * tex + 0.1 * torch.sin(hit[..., 2]), 1.0 - tex], -1).clamp(0, 1).float()
        poses[s] = torch.eye(4)
        poses[s, :3, :3], poses[s, :3, 3] = R.float(), pos.float()
        pos = pos + torch.tensor([0.02, 0.003, 0.015], dtype=torch.float64) + 0.002 * torch.randn(3, generator=g, dtype=torch.float64)
        yaw += math.radians(0.6)
        pitch += math.radians(0.1) * math.sin(s / 5.0)
    return depth, rgb, K.to(device), poses.to(device)
