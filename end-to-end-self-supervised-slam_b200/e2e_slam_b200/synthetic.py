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
