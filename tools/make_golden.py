#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules on CPU.

Runs only in the build container, where /root/reference exists (the GPU box does
not have it; tests read the committed .npz files instead).  The reference imports
`matplotlib` and `chamferdist` at module scope (loss/losses.py:3-4,
depth_estimation/networks.py:14); neither is installed, neither is used by the
functions exercised here, so empty stub modules are registered first.

What is recorded per case (all produced by reference code, fp32 unless noted):
  inputs      depth, inv_K, K, T, colors (B,2,H,W,3 channels-last like gradslam's loader)
  forward     cam_points, pix_coords, valid_mask, synthesized frame, SSIM map,
              per-pixel photometric loss map, scalar loss (mean)
  backward    d loss / d depth, d loss / d source image, d loss / d T   (fp32 autograd)
  fp64        the same forward scalars and gradients evaluated in float64 ("truth")
  extras      smoothness / sparse-gt / regulariser / geometric losses (+ gradients)

Usage:  python tools/make_golden.py            (writes tests/golden/)
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("E2E_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "chamferdist", "chamferdist.chamfer"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["chamferdist.chamfer"].knn_points = None
    sys.modules["chamferdist"].ChamferDistance = None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    from depth_estimation.view_synthesis import BackprojectDepth, Project3D  # noqa
    from loss import losses  # noqa
    return BackprojectDepth, Project3D, losses


def so3_exp(w):
    th = float(np.linalg.norm(w))
    if th < 1e-12:
        return np.eye(3)
    k = w / th
    kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * kx + (1 - np.cos(th)) * kx @ kx


def make_case(seed, B, H, W, kind, rot_deg, trans, holes=False):
    """Synthetic ICL/TUM-shaped inputs, scaled to H x W."""
    rng = np.random.default_rng(seed)
    if kind == "icl":   # ICL-NUIM intrinsics, note the negative fy
        fx, fy, cx, cy = 481.2 * W / 640, -480.0 * H / 480, W / 2 - 0.5, H / 2 - 0.5
    else:               # TUM
        fx, fy, cx, cy = 525.0 * W / 640, 525.0 * H / 480, W / 2 - 0.5, H / 2 - 0.5
    K = np.eye(4, dtype=np.float64)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2] = fx, fy, cx, cy
    K = np.repeat(K[None], B, 0)
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    depth = np.empty((B, 1, H, W))
    colors = np.empty((B, 2, H, W, 3))
    T = np.empty((B, 4, 4))
    for b in range(B):
        a = rng.uniform(-0.5, 0.5, 2)
        depth[b, 0] = 2.0 + a[0] * xx / W + a[1] * yy / H + 0.3 * np.sin(xx / 5.0 + b) * np.cos(yy / 4.0)
        depth[b, 0] += rng.uniform(0, 0.05, (H, W))
        for f in range(2):
            for c in range(3):
                img = np.zeros((H, W))
                for _ in range(8):
                    kx_, ky_ = rng.uniform(-0.6, 0.6, 2)
                    img += rng.uniform(0.2, 1.0) * np.sin(kx_ * xx + ky_ * yy + rng.uniform(0, 6.28))
                img = (img - img.min()) / (img.max() - img.min() + 1e-9)
                colors[b, f, :, :, c] = 0.96 * img + rng.uniform(0, 0.04, (H, W))
        w = rng.normal(size=3)
        w = w / np.linalg.norm(w) * np.deg2rad(rot_deg) * rng.uniform(0.5, 1.0)
        t = rng.normal(size=3)
        t = t / np.linalg.norm(t) * trans * rng.uniform(0.5, 1.0)
        T[b] = np.eye(4)
        T[b, :3, :3] = so3_exp(w)
        T[b, :3, 3] = t
    if holes:  # saturated / flat regions: exact zeros and ones in the images
        colors[:, :, : H // 4, : W // 3, :] = 0.0
        colors[:, :, -H // 5:, -W // 4:, :] = 1.0
    return (torch.tensor(depth, dtype=torch.float32), torch.tensor(K, dtype=torch.float32),
            torch.tensor(T, dtype=torch.float32), torch.tensor(colors, dtype=torch.float32))


def run_reference(BackprojectDepth, Project3D, losses, depth, K, T, colors, padding_mode, use_mask, dtype):
    """The reference's own call sequence: train_depth.py:545-613 (novel_view_synthesis),
    :707-727 (compute_photometric_loss), :629/:657 (mean over frames, mean over pixels)."""
    B, _, H, W = depth.shape
    depth = depth.to(dtype).clone().requires_grad_(True)
    K = K.to(dtype)
    T = T.to(dtype).clone().requires_grad_(True)
    colors = colors.to(dtype).clone().requires_grad_(True)
    inv_K = torch.pinverse(K.float()).to(dtype)       # train_depth.py:460-461 (always fp32 in the reference)
    src = colors[:, 0].permute(0, 3, 1, 2)            # train_depth.py:451-453: NCHW *view* of NHWC memory
    tgt = colors[:, 1].permute(0, 3, 1, 2)
    bp = BackprojectDepth(B, H, W).to(dtype)
    pj = Project3D(B, H, W).to(dtype)
    ssim = losses.SSIM()
    cam = bp(depth, inv_K)
    pix, valid = pj(cam, K, T, False)
    syn = F.grid_sample(src, pix, padding_mode=padding_mode, align_corners=False)
    if use_mask:
        pred, target = syn * valid, tgt * valid
    else:
        pred, target = syn, tgt
    ssim_map = ssim(pred, target)
    loss_map = losses.photometric_loss(ssim, pred, target)
    loss = loss_map.mean(1, keepdim=True).mean()
    loss.backward()
    g_src = colors.grad[:, 0]   # channels-last (B,H,W,3)
    return dict(inv_K=inv_K, cam=cam, pix=pix, valid=valid, syn=syn, ssim=ssim_map, loss_map=loss_map,
                loss=loss, g_depth=depth.grad, g_src=g_src, g_T=T.grad)


def np32(t):
    return t.detach().contiguous().cpu().numpy()


def main():
    os.makedirs(OUT, exist_ok=True)
    BackprojectDepth, Project3D, losses = _import_reference()
    torch.set_num_threads(1)   # fix the reduction tree of torch's own .mean() for the scalar
    cases = [
        # name,           seed B  H   W   kind  rot  trans pad      mask  holes
        ("icl_border",      0, 1, 24, 32, "icl", 2.0, 0.05, "border", True, False),
        ("tum_zeros",       1, 1, 20, 28, "tum", 5.0, 0.40, "zeros",  True, False),
        ("icl_b2_nomask",   2, 2, 32, 40, "icl", 3.0, 0.10, "border", False, False),
        ("tum_flat",        3, 1, 36, 52, "tum", 4.0, 0.15, "border", True, True),
        ("icl_ragged",      4, 3, 19, 67, "icl", 6.0, 0.60, "zeros",  True, False),
    ]
    for name, seed, B, H, W, kind, rot, trans, pad, mask, holes in cases:
        depth, K, T, colors = make_case(seed, B, H, W, kind, rot, trans, holes)
        r32 = run_reference(BackprojectDepth, Project3D, losses, depth, K, T, colors, pad, mask, torch.float32)
        r64 = run_reference(BackprojectDepth, Project3D, losses, depth, K, T, colors, pad, mask, torch.float64)
        out = dict(depth=np32(depth), K=np32(K), T=np32(T), colors=np32(colors),
                   padding_mode=np.array(pad), use_mask=np.array(mask))
        for k, v in r32.items():
            out[k] = np32(v)
        for k in ("loss", "loss_map", "g_depth", "g_src", "g_T", "syn"):
            out[k + "_f64"] = np32(r64[k])
        # ---- extras: smoothness (train_depth.py:763-773 + losses.py:119-132), sparse gt (losses.py:151-160),
        # regulariser (losses.py:134-148) -------------------------------------------------------------------
        g = torch.Generator().manual_seed(seed)
        disp = (torch.rand(B, 1, H, W, generator=g) * 2 + 0.05).requires_grad_(True)
        tgt = colors[:, 1].permute(0, 3, 1, 2)
        norm = disp / (disp.mean(2, True).mean(3, True) + 1e-7)
        sm = losses.disparity_smoothness_loss(norm, tgt)
        sm.backward()
        out.update(disp=np32(disp), smooth=np32(sm), g_disp_smooth=np32(disp.grad))
        if B == 1:  # depth_gt_loss's .squeeze() broadcasting is only meaningful for B == 1
            gt = depth.permute(0, 2, 3, 1) * (torch.rand(B, H, W, 1, generator=g) > 0.15)
            m = (torch.rand(B, H, W, 1, generator=g) < 0.3).float() * (gt != 0)
            pred = (depth * 1.1).clone().requires_grad_(True)
            l = losses.depth_gt_loss(pred, gt * m, m)
            l.backward()
            out.update(sparse_gt=np32(gt * m), sparse_mask=np32(m), pred_depth=np32(pred),
                       gt_loss=np32(l), g_pred_gt=np32(pred.grad))
        for kind_ in ("l1", "l2"):
            a = depth.clone()
            b_ = (depth * 1.05 + 0.01).clone().requires_grad_(True)
            l = losses.depth_reguralizer(a, b_, kind_)
            l.backward()
            out["reg_" + kind_] = np32(l)
            out["g_reg_" + kind_] = np32(b_.grad)
        # geometric consistency (losses.py:84-95) with Project3D(geometric=True) (view_synthesis.py:73-76)
        d2 = depth.clone().requires_grad_(True)
        inv_K = torch.pinverse(K)
        cam = BackprojectDepth(B, H, W)(d2, inv_K)
        pix, wdepth, valid = Project3D(B, H, W)(cam, K, T, True)
        src_depth = depth * 1.03
        idepth = F.grid_sample(src_depth, pix, padding_mode=pad, align_corners=False)
        o = {("warped_depth", -1): wdepth, ("interpolated_depth", -1): idepth, ("valid_mask", -1): valid}
        old = losses.geometric_consistency_loss.__defaults__
        gl = losses.geometric_consistency_loss(o, -1, torch.device("cpu"))
        out.update(warped_depth=np32(wdepth), interp_depth=np32(idepth), geo_loss=np32(gl),
                   src_depth=np32(src_depth))
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: loss={float(r32['loss']):.8f} (f64 {float(r64['loss']):.10f}) valid={float(r32['valid'].mean()):.3f} "
              f"-> {os.path.getsize(path)/1024:.0f} KiB")


if __name__ == "__main__":
    main()
