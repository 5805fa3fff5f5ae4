#!/usr/bin/env python
"""Generate tests/golden/composite_*.npz: the reference's multi-source photometric objective -- mean over source frames,
min-reprojection and auto-masking (train_depth.py:615-660, 707-750) -- evaluated with the UNMODIFIED reference modules
(BackprojectDepth, Project3D, SSIM, photometric_loss imported from /root/reference; F.grid_sample) on CPU.

`compute_losses` itself is a method of the driver class (train_depth.py:615), which cannot be imported here (it pulls in gradslam
and the datasets), so the ~20 lines that combine the per-frame maps (`.mean(1, keepdim=True)`, `torch.cat((auto_masking,
photometric), 1)`, `torch.min(dim=1)`, `.mean()`) are written out below exactly as they stand there, around the reference's own
loss modules.  The tie-breaking noise (`torch.randn(...) * 0.00001`, :646) is drawn once from a seeded generator and stored, so
that a test can feed the same noise.

Usage:  python tools/make_golden_composite.py        (build container only; writes tests/golden/composite_*.npz)
"""
import importlib.util
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)


def objective(BackprojectDepth, Project3D, losses, depth, K, Ts, colors, padding_mode, use_mask, min_reprojection, auto_masking,
              noise, dtype):
    """colors (B, 1 + S, H, W, 3): frame 0 = target, frames 1.. = sources; Ts list of (B,4,4)."""
    B, _, H, W = depth.shape
    depth = depth.to(dtype).clone().requires_grad_(True)
    K = K.to(dtype)
    colors = colors.to(dtype).clone().requires_grad_(True)
    inv_K = torch.pinverse(K.float()).to(dtype)
    tgt = colors[:, 0].permute(0, 3, 1, 2)
    bp, pj, ssim = BackprojectDepth(B, H, W).to(dtype), Project3D(B, H, W).to(dtype), losses.SSIM()
    photometric_losses, auto_masking_losses = [], []
    cam = bp(depth, inv_K)                                                           # train_depth.py:560
    for s, T in enumerate(Ts):
        src = colors[:, 1 + s].permute(0, 3, 1, 2)
        pix, valid = pj(cam, K, T.to(dtype), False)                                  # :578-580
        syn = F.grid_sample(src, pix, padding_mode=padding_mode, align_corners=False)    # :587-590
        if use_mask:                                                                 # :711-718, 736-742
            photometric_losses.append(losses.photometric_loss(ssim, syn * valid, tgt * valid))
            auto_masking_losses.append(losses.photometric_loss(ssim, src * valid, tgt * valid))
        else:
            photometric_losses.append(losses.photometric_loss(ssim, syn, tgt))
            auto_masking_losses.append(losses.photometric_loss(ssim, src, tgt))
    photmetric = torch.cat(photometric_losses, 1)                                    # :726
    if not min_reprojection:                                                         # :624-629
        photmetric = photmetric.mean(1, keepdim=True)
    if auto_masking:                                                                 # :642-651
        am = torch.cat(auto_masking_losses, 1)                                       # :749
        if min_reprojection:
            am = am + noise.to(dtype)                                                # :646 (the noise tensor is supplied)
        else:
            am = am.mean(1, keepdim=True)
        photmetric = torch.cat((am, photmetric), dim=1)
    if photmetric.shape[1] == 1:                                                     # :653-658
        optimize = photmetric.mean()
        index = torch.zeros(B, H, W, dtype=torch.int64)
    else:
        optimize, index = torch.min(photmetric, dim=1)
        optimize = optimize.mean()
    optimize.backward()
    return dict(loss=optimize.detach(), index=index, g_depth=depth.grad, g_colors=colors.grad)


def main():
    BackprojectDepth, Project3D, losses = mg._import_reference()
    torch.set_num_threads(1)
    cases = [("composite_icl_border", 10, 1, 24, 32, "icl", 3.0, 0.10, "border", True),
             ("composite_tum_zeros_b2", 11, 2, 26, 36, "tum", 5.0, 0.30, "zeros", True),
             ("composite_icl_nomask", 12, 1, 21, 40, "icl", 2.0, 0.05, "border", False)]
    for name, seed, B, H, W, kind, rot, trans, pad, mask in cases:
        depth, K, T1, c01 = mg.make_case(seed, B, H, W, kind, rot, trans)
        _, _, T2, c23 = mg.make_case(seed + 100, B, H, W, kind, rot, trans)
        colors = torch.cat([c01[:, 1:2], c01[:, 0:1], c23[:, 0:1]], 1)                 # target, source -1, source +1
        # make the sources resemble the target so that warped and identity candidates really compete
        colors[:, 1:] = 0.6 * colors[:, 0:1] + 0.4 * colors[:, 1:]
        Ts = [T1, torch.linalg.inv(T2)]
        noise = torch.randn(B, 2, H, W, generator=torch.Generator().manual_seed(seed)) * 0.00001
        out = dict(depth=mg.np32(depth), K=mg.np32(K), inv_K=mg.np32(torch.pinverse(K.float())), T=np.stack([mg.np32(t) for t in Ts], 1), colors=mg.np32(colors), noise=mg.np32(noise),
                   padding_mode=np.array(pad), use_mask=np.array(mask))
        for mr in (False, True):
            for am in (False, True):
                tag = f"mr{int(mr)}_am{int(am)}"
                r32 = objective(BackprojectDepth, Project3D, losses, depth, K, Ts, colors, pad, mask, mr, am, noise, torch.float32)
                r64 = objective(BackprojectDepth, Project3D, losses, depth, K, Ts, colors, pad, mask, mr, am, noise, torch.float64)
                out[f"loss_{tag}"], out[f"index_{tag}"] = mg.np32(r32["loss"]), r32["index"].numpy().astype(np.int8)
                out[f"g_depth_{tag}"], out[f"g_colors_{tag}"] = mg.np32(r32["g_depth"]), mg.np32(r32["g_colors"])
                out[f"loss_{tag}_f64"], out[f"g_depth_{tag}_f64"] = mg.np32(r64["loss"]), mg.np32(r64["g_depth"])
                out[f"g_colors_{tag}_f64"] = mg.np32(r64["g_colors"])
                sel = np.bincount(r32["index"].numpy().ravel(), minlength=4)
                print(f"{name} {tag}: loss {float(r32['loss']):.8f}  candidates chosen {sel.tolist()}")
        path = os.path.join(mg.OUT, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"  -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
