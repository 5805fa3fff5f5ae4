"""Drop-in for the reference's depth_estimation/view_synthesis.py: same class names, constructor
arguments, call signatures and return values, backed by one CUDA kernel per call (tier (i)).

Differences that are deliberate and invisible to the reference's scripts:
  * no per-(B,H,W) pixel-grid buffers are allocated (view_synthesis.py:17-32): the kernels derive x, y
    from the thread index, so any batch size is accepted at call time;
  * inputs must be CUDA float32 tensors; there is no CPU path.
"""
import ctypes

import torch
import torch.nn as nn

from ._lib import check, f32, lib, prepare_divisors, ptr, stream_ptr, strides4

_PAD = {"zeros": 0, "border": 1}


def _k44(m, B, name):
    f32(m, name)
    if m.dim() != 3 or m.shape[1:] != (4, 4):
        raise ValueError(f"{name} must be (B,4,4), got {tuple(m.shape)}")
    if m.shape[0] == 1 and B > 1:
        m = m.expand(B, 4, 4)
    if m.shape[0] != B:
        raise ValueError(f"{name} batch {m.shape[0]} does not match {B}")
    return m.contiguous()


class _Backproject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, inv_K):
        f32(depth, "depth")
        B, _, H, W = depth.shape
        depth_c, inv_K_c = depth.contiguous(), _k44(inv_K, B, "inv_K")
        cam = torch.empty(B, 4, H * W, dtype=torch.float32, device=depth.device)
        with torch.cuda.device(depth.device):
            check(lib().e2e_backproject_fwd(ptr(depth_c), ptr(inv_K_c), B, H, W, ptr(cam), stream_ptr()), "e2e_backproject_fwd")
        ctx.save_for_backward(inv_K_c)
        ctx.shape = (B, H, W)
        return cam

    @staticmethod
    def backward(ctx, g):
        (inv_K,) = ctx.saved_tensors
        B, H, W = ctx.shape
        g = f32(g, "grad").contiguous()
        gd = torch.empty(B, 1, H, W, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            check(lib().e2e_backproject_bwd(ptr(g), ptr(inv_K), B, H, W, ptr(gd), stream_ptr()), "e2e_backproject_bwd")
        return gd, None


class BackprojectDepth(nn.Module):
    """Transform a depth map into a homogeneous point cloud (view_synthesis.py:7-40).
    forward(depth [B,1,H,W], inv_K [B,4,4]) -> cam_points [B,4,H*W]."""

    def __init__(self, batch_size, height, width):
        super().__init__()
        self.batch_size, self.height, self.width = batch_size, height, width

    def forward(self, depth, inv_K):
        if depth.dim() != 4 or depth.shape[1] != 1 or depth.shape[2:] != (self.height, self.width):
            raise ValueError(f"depth must be (B,1,{self.height},{self.width}), got {tuple(depth.shape)}")
        return _Backproject.apply(depth, inv_K)


class _Project(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, K, T, H, W, eps, geometric):
        f32(points, "points")
        B = points.shape[0]
        if points.dim() != 3 or points.shape[1] != 4 or points.shape[2] != H * W:
            raise ValueError(f"points must be (B,4,{H * W}), got {tuple(points.shape)}")
        pts, K_c, T_c = points.contiguous(), _k44(K, B, "K"), _k44(T, B, "T")
        prepare_divisors(W - 1, H - 1)
        dev = points.device
        pix = torch.empty(B, H, W, 2, dtype=torch.float32, device=dev)
        valid = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev)
        wdepth = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev) if geometric else None
        with torch.cuda.device(dev):
            check(lib().e2e_project3d_fwd(ptr(pts), ptr(K_c), ptr(T_c), B, H, W, ctypes.c_float(eps), ptr(pix), ptr(valid),
                                          ptr(wdepth), stream_ptr()), "e2e_project3d_fwd")
        ctx.save_for_backward(pts, K_c, T_c)
        ctx.cfg = (B, H, W, eps, geometric)
        ctx.mark_non_differentiable(valid)
        return (pix, wdepth, valid) if geometric else (pix, valid)

    @staticmethod
    def backward(ctx, *grads):
        pts, K, T = ctx.saved_tensors
        B, H, W, eps, geometric = ctx.cfg
        g_pix = grads[0]
        g_wd = grads[1] if geometric else None
        dev = pts.device
        if g_pix is None and g_wd is None:
            return (None,) * 7
        g_pix = f32(g_pix, "grad").contiguous() if g_pix is not None else None
        g_wd = f32(g_wd, "grad").contiguous() if g_wd is not None else None
        g_pts = torch.empty_like(pts)
        need_P = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        g_P = torch.empty(B, 3, 4, dtype=torch.float32, device=dev) if need_P else None
        nws = 12 * 4 * 296 * B + 256
        ws = torch.empty(nws, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib().e2e_project3d_bwd(ptr(pts), ptr(K), ptr(T), B, H, W, ctypes.c_float(eps), ptr(g_pix), ptr(g_wd),
                                          ptr(g_pts), ptr(g_P), ptr(ws), nws, stream_ptr()), "e2e_project3d_bwd")
        g_K = g_T = None
        if g_P is not None:
            if ctx.needs_input_grad[2]:
                g_T = torch.matmul(K[:, :3, :].transpose(1, 2), g_P)
            if ctx.needs_input_grad[1]:
                g_K = torch.zeros_like(K)
                g_K[:, :3, :] = torch.matmul(g_P, T.transpose(1, 2))
        return g_pts, g_K, g_T, None, None, None, None


class Project3D(nn.Module):
    """Project 3D points into a camera with intrinsics K at pose T (view_synthesis.py:42-78).
    forward(points [B,4,HW], K, T, geometric) -> (pix [B,H,W,2], valid [B,1,H,W]) or, with geometric=True,
    (pix, warped_depth [B,1,H,W], valid)."""

    def __init__(self, batch_size, height, width, eps=1e-7):
        super().__init__()
        self.batch_size, self.height, self.width, self.eps = batch_size, height, width, eps

    def forward(self, points, K, T, geometric):
        return _Project.apply(points, K, T, self.height, self.width, float(self.eps), bool(geometric))


class _GridSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, grid, pad, align):
        f32(inp, "input"), f32(grid, "grid")
        if inp.dim() != 4 or grid.dim() != 4 or grid.shape[-1] != 2 or grid.shape[0] != inp.shape[0]:
            raise ValueError(f"expected input (B,C,H,W) and grid (B,Ho,Wo,2), got {tuple(inp.shape)} / {tuple(grid.shape)}")
        B, C, H, W = inp.shape
        Ho, Wo = grid.shape[1:3]
        grid_c = grid.contiguous()
        out = torch.empty(B, C, Ho, Wo, dtype=torch.float32, device=inp.device)
        with torch.cuda.device(inp.device):
            check(lib().e2e_grid_sample_fwd(ptr(inp), strides4(inp), ptr(grid_c), B, C, H, W, Ho, Wo, pad, align, ptr(out),
                                            stream_ptr()), "e2e_grid_sample_fwd")
        ctx.save_for_backward(inp, grid_c)
        ctx.cfg = (pad, align)
        return out

    @staticmethod
    def backward(ctx, g):
        inp, grid = ctx.saved_tensors
        pad, align = ctx.cfg
        B, C, H, W = inp.shape
        Ho, Wo = grid.shape[1:3]
        g = f32(g, "grad").contiguous()
        g_in = torch.zeros(B, C, H, W, dtype=torch.float32, device=inp.device) if ctx.needs_input_grad[0] else None
        g_grid = torch.empty_like(grid) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(inp.device):
            check(lib().e2e_grid_sample_bwd(ptr(g), ptr(inp), strides4(inp), ptr(grid), B, C, H, W, Ho, Wo, pad, align,
                                            ptr(g_in), strides4(g_in) if g_in is not None else None, ptr(g_grid),
                                            stream_ptr()), "e2e_grid_sample_bwd")
        return g_in, g_grid, None, None


def grid_sample(input, grid, mode="bilinear", padding_mode="zeros", align_corners=False):
    """torch.nn.functional.grid_sample for the configurations the reference uses (train_depth.py:568-590):
    bilinear, padding 'zeros' | 'border', align_corners False | True.  `patch.install()` can route a
    script's F.grid_sample calls here."""
    if mode != "bilinear":
        raise NotImplementedError("only bilinear sampling is used by the reference")
    if padding_mode not in _PAD:
        raise ValueError(f"padding_mode must be 'zeros' or 'border', got {padding_mode!r}")
    return _GridSample.apply(input, grid, _PAD[padding_mode], int(bool(align_corners)))
