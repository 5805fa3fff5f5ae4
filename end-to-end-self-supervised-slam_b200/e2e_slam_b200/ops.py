"""torch.autograd bindings of the fused CUDA kernels (tier (ii) of SURVEY.md section 8(b)).

`warp_photometric` / `warp_photometric_loss` replace, for one source frame, the reference's
    BackprojectDepth.forward -> Project3D.forward -> F.grid_sample -> mask multiply -> SSIM ->
    photometric_loss (-> .mean())
(train_depth.py:545-613 + 707-727; online_adaption.py:412-455 + 544-564) with one forward and one
backward kernel.  Everything is fp32; all tensors must live on a CUDA device.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import check, f32, lib, prepare_divisors, ptr, stream_ptr, strides4

_PAD = {"zeros": 0, "border": 1}


def _pad_code(padding_mode):
    if padding_mode not in _PAD:
        raise ValueError(f"padding_mode must be 'zeros' or 'border' (MODEL.padding_mode), got {padding_mode!r}")
    return _PAD[padding_mode]


def _mat44(m, B, name):
    f32(m, name)
    if m.dim() == 2:
        m = m.unsqueeze(0)
    if m.shape[-2:] != (4, 4):
        raise ValueError(f"{name} must be (B,4,4), got {tuple(m.shape)}")
    if m.shape[0] != B:
        if m.shape[0] != 1:
            raise ValueError(f"{name} batch {m.shape[0]} does not match depth batch {B}")
        m = m.expand(B, 4, 4)
    return m.contiguous()


def _workspace(B, H, W, device):
    n = lib().e2e_warp_photo_workspace_bytes(B, H, W)
    return torch.empty(n, dtype=torch.uint8, device=device), n


_SPECULATE = os.environ.get("E2E_SPECULATE", "1") != "0"


class _WarpPhotometric(torch.autograd.Function):
    """mode 'map'  -> returns loss_map [B,1,H,W] (+ syn, valid, pix when materialise=True)
       mode 'mean' -> returns the scalar mean of the loss map (lean path, nothing else is written)

    Map mode with autograd: the reference reduces the map with `.mean(1, keepdim=True).mean()` (train_depth.py:629, 657), so
    the gradient that comes back is the same number at every pixel.  forward() therefore runs the single-sweep kernel
    (e2e_warp_photo_vg_map), which writes the map AND the gradients for that uniform case; backward() checks the upstream
    map on the device (e2e_upstream_uniform), rescales the stored gradients, and enqueues the streaming backward kernel
    with a skip flag -- it only does work when the upstream gradient is not uniform (min-reprojection, weighting).  No host
    synchronisation either way.  E2E_SPECULATE=0 restores forward kernel + unconditional backward kernel."""

    @staticmethod
    def forward(ctx, depth, inv_K, K, T, src, tgt, padding_mode, use_mask, eps, mode, materialise, expect_uniform=True):
        f32(depth, "depth"), f32(src, "source frame"), f32(tgt, "target frame")
        if depth.dim() != 4 or depth.shape[1] != 1:
            raise ValueError(f"depth must be (B,1,H,W), got {tuple(depth.shape)}")
        B, _, H, W = depth.shape
        if tuple(src.shape) != (B, 3, H, W) or tuple(tgt.shape) != (B, 3, H, W):
            raise ValueError(f"source/target frames must be ({B},3,{H},{W}), got {tuple(src.shape)} / {tuple(tgt.shape)}")
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("the fused op does not differentiate inv_K (the reference derives it from the dataset's K with "
                                      "torch.pinverse outside the graph, train_depth.py:460-461); detach it")
        depth_c = depth.contiguous()
        inv_K_c, K_c, T_c = _mat44(inv_K, B, "inv_K"), _mat44(K, B, "K"), _mat44(T, B, "T")
        prepare_divisors(W - 1, H - 1, 9.0, 3.0)
        dev = depth.device
        pad = _pad_code(padding_mode)
        ws, ws_bytes = _workspace(B, H, W, dev)
        syn = valid = pix = loss_map = loss_mean = None
        need_depth, _, need_K, need_T, need_src = ctx.needs_input_grad[:5]
        ctx.spec = None
        if mode == "map" and _SPECULATE and expect_uniform and (need_depth or need_K or need_T or need_src):
            loss_map = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev)
            if materialise:
                syn = torch.empty(B, 3, H, W, dtype=torch.float32, device=dev)
                valid = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev)
                pix = torch.empty(B, H, W, 2, dtype=torch.float32, device=dev)
            grad_depth = torch.empty_like(depth_c)
            grad_src = torch.zeros(B, 3, H, W, dtype=torch.float32, device=dev) if need_src else None
            grad_P = torch.empty(B, 3, 4, dtype=torch.float32, device=dev) if (need_K or need_T) else None
            with torch.cuda.device(dev):
                rc = lib().e2e_warp_photo_vg_map(ptr(depth_c), ptr(inv_K_c), ptr(K_c), ptr(T_c), ptr(src), strides4(src),
                                                 ptr(tgt), strides4(tgt), B, H, W, pad, int(bool(use_mask)), ctypes.c_float(eps),
                                                 ptr(loss_map), ptr(syn), ptr(valid), ptr(pix), None,
                                                 ptr(grad_depth), ptr(grad_src),
                                                 strides4(grad_src) if grad_src is not None else None,
                                                 ptr(grad_P), ptr(ws), ws_bytes, stream_ptr())
            check(rc, "e2e_warp_photo_vg_map")
            ctx.spec = (grad_depth, grad_src, grad_P)
            ctx.save_for_backward(depth_c, inv_K_c, K_c, T_c, src, tgt)
            ctx.cfg = (B, H, W, pad, int(bool(use_mask)), float(eps), mode)
            if materialise:
                ctx.mark_non_differentiable(syn, valid, pix)
                return loss_map, syn, valid, pix
            return loss_map
        if mode == "map":
            loss_map = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev)
            if materialise:
                syn = torch.empty(B, 3, H, W, dtype=torch.float32, device=dev)
                valid = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev)
                pix = torch.empty(B, H, W, 2, dtype=torch.float32, device=dev)
        else:
            loss_mean = torch.empty(1, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib().e2e_warp_photo_fwd(ptr(depth_c), ptr(inv_K_c), ptr(K_c), ptr(T_c),
                                          ptr(src), strides4(src), ptr(tgt), strides4(tgt),
                                          B, H, W, pad, int(bool(use_mask)), ctypes.c_float(eps),
                                          ptr(syn), ptr(valid), ptr(pix), ptr(loss_map), ptr(loss_mean),
                                          ptr(ws), ws_bytes, stream_ptr())
        check(rc, "e2e_warp_photo_fwd")
        ctx.save_for_backward(depth_c, inv_K_c, K_c, T_c, src, tgt)
        ctx.cfg = (B, H, W, pad, int(bool(use_mask)), float(eps), mode)
        if mode == "map":
            if materialise:
                ctx.mark_non_differentiable(syn, valid, pix)
                return loss_map, syn, valid, pix
            return loss_map
        return loss_mean.reshape(())

    @staticmethod
    def backward(ctx, *grads):
        depth, inv_K, K, T, src, tgt = ctx.saved_tensors
        B, H, W, pad, use_mask, eps, mode = ctx.cfg
        g = grads[0]
        dev = depth.device
        need_depth, _, need_K, need_T, need_src = ctx.needs_input_grad[:5]
        ws, ws_bytes = _workspace(B, H, W, dev)
        if ctx.spec is not None:                      # gradients for a uniform upstream gradient were produced by forward()
            grad_depth, grad_src, grad_P = ctx.spec
            ctx.spec = None
            g_map = f32(g, "grad").contiguous()
            scale2 = torch.empty(2, dtype=torch.float32, device=dev)
            n = B * H * W
            with torch.cuda.device(dev):
                check(lib().e2e_upstream_uniform(ptr(g_map), n, float(n), ptr(scale2), stream_ptr()), "e2e_upstream_uniform")
                check(lib().e2e_scale_or_zero(ptr(grad_depth), grad_depth.numel(), ptr(grad_src),
                                              grad_src.numel() if grad_src is not None else 0, ptr(grad_P),
                                              grad_P.numel() if grad_P is not None else 0, ptr(scale2), stream_ptr()),
                      "e2e_scale_or_zero")
                rc = lib().e2e_warp_photo_bwd_cond(ptr(depth), ptr(inv_K), ptr(K), ptr(T), ptr(src), strides4(src), ptr(tgt),
                                                   strides4(tgt), B, H, W, pad, use_mask, ctypes.c_float(eps), ptr(g_map), None,
                                                   ctypes.c_float(1.0), ptr(scale2[1:]), ptr(grad_depth), ptr(grad_src),
                                                   strides4(grad_src) if grad_src is not None else None,
                                                   ptr(grad_P), ptr(ws), ws_bytes, stream_ptr())
            check(rc, "e2e_warp_photo_bwd_cond")
            return _WarpPhotometric._finish(ctx, K, T, grad_depth, grad_src, grad_P)
        grad_depth = torch.empty_like(depth)
        grad_src = torch.zeros(B, 3, H, W, dtype=torch.float32, device=dev) if need_src else None
        grad_P = torch.empty(B, 3, 4, dtype=torch.float32, device=dev) if (need_K or need_T) else None
        if mode == "map":
            g_map, g_scalar, scale = f32(g, "grad").contiguous(), None, 1.0
        else:
            g_map, g_scalar, scale = None, f32(g, "grad").reshape(1).contiguous(), 1.0 / (B * H * W)
        with torch.cuda.device(dev):
            rc = lib().e2e_warp_photo_bwd(ptr(depth), ptr(inv_K), ptr(K), ptr(T),
                                          ptr(src), strides4(src), ptr(tgt), strides4(tgt),
                                          B, H, W, pad, use_mask, ctypes.c_float(eps),
                                          ptr(g_map), ptr(g_scalar), ctypes.c_float(scale),
                                          ptr(grad_depth), ptr(grad_src),
                                          strides4(grad_src) if grad_src is not None else None,
                                          ptr(grad_P), ptr(ws), ws_bytes, stream_ptr())
        check(rc, "e2e_warp_photo_bwd")
        return _WarpPhotometric._finish(ctx, K, T, grad_depth, grad_src, grad_P)

    @staticmethod
    def _finish(ctx, K, T, grad_depth, grad_src, grad_P):
        need_depth, _, need_K, need_T, need_src = ctx.needs_input_grad[:5]
        grad_K = grad_T = None
        if grad_P is not None:
            # P = (K @ T)[:3]  =>  dL/dT = K[:3]^T dL/dP ,  dL/dK[:3] = dL/dP T^T   (4x4 host-side plumbing)
            if need_T:
                grad_T = torch.matmul(K[:, :3, :].transpose(1, 2), grad_P)
            if need_K:
                grad_K = torch.zeros_like(K)
                grad_K[:, :3, :] = torch.matmul(grad_P, T.transpose(1, 2))
        return (grad_depth if need_depth else None, None, grad_K, grad_T, grad_src, None,
                None, None, None, None, None, None)


class _WarpPhotometricMean(torch.autograd.Function):
    """Scalar mean photometric loss through the single-pass value+gradient kernel (e2e_warp_photo_vg):
    forward() produces the loss and, in the same sweep, its gradients for an upstream gradient of 1;
    backward() only rescales them by the actual upstream scalar (a kernel that exits immediately when that
    scalar is 1, which is the reference's `loss.backward()`)."""

    @staticmethod
    def forward(ctx, depth, inv_K, K, T, src, tgt, padding_mode, use_mask, eps):
        f32(depth, "depth"), f32(src, "source frame"), f32(tgt, "target frame")
        if depth.dim() != 4 or depth.shape[1] != 1:
            raise ValueError(f"depth must be (B,1,H,W), got {tuple(depth.shape)}")
        B, _, H, W = depth.shape
        if tuple(src.shape) != (B, 3, H, W) or tuple(tgt.shape) != (B, 3, H, W):
            raise ValueError(f"source/target frames must be ({B},3,{H},{W}), got {tuple(src.shape)} / {tuple(tgt.shape)}")
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("the fused op does not differentiate inv_K (the reference derives it from the dataset's K with "
                                      "torch.pinverse outside the graph, train_depth.py:460-461); detach it")
        depth_c = depth.contiguous()
        inv_K_c, K_c, T_c = _mat44(inv_K, B, "inv_K"), _mat44(K, B, "K"), _mat44(T, B, "T")
        prepare_divisors(W - 1, H - 1, 9.0, 3.0)
        dev = depth.device
        pad = _pad_code(padding_mode)
        need_depth, _, need_K, need_T, need_src = ctx.needs_input_grad[:5]
        if not (need_depth or need_K or need_T or need_src):       # value only: the lean forward kernel
            ws, ws_bytes = _workspace(B, H, W, dev)
            loss = torch.empty(1, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                rc = lib().e2e_warp_photo_fwd(ptr(depth_c), ptr(inv_K_c), ptr(K_c), ptr(T_c), ptr(src), strides4(src),
                                              ptr(tgt), strides4(tgt), B, H, W, pad, int(bool(use_mask)),
                                              ctypes.c_float(eps), None, None, None, None, ptr(loss), ptr(ws), ws_bytes,
                                              stream_ptr())
            check(rc, "e2e_warp_photo_fwd")
            return loss.reshape(())
        n = lib().e2e_warp_photo_vg_workspace_bytes(B, H, W)
        ws = torch.empty(n, dtype=torch.uint8, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        grad_depth = torch.empty_like(depth_c)
        grad_src = torch.zeros(B, 3, H, W, dtype=torch.float32, device=dev) if need_src else None
        grad_P = torch.empty(B, 3, 4, dtype=torch.float32, device=dev) if (need_K or need_T) else None
        with torch.cuda.device(dev):
            rc = lib().e2e_warp_photo_vg(ptr(depth_c), ptr(inv_K_c), ptr(K_c), ptr(T_c), ptr(src), strides4(src),
                                         ptr(tgt), strides4(tgt), B, H, W, pad, int(bool(use_mask)), ctypes.c_float(eps),
                                         ptr(loss), ptr(grad_depth), ptr(grad_src),
                                         strides4(grad_src) if grad_src is not None else None,
                                         ptr(grad_P), ptr(ws), n, stream_ptr())
        check(rc, "e2e_warp_photo_vg")
        ctx.grads = (grad_depth, grad_src, grad_P)
        ctx.save_for_backward(K_c, T_c)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        if ctx.grads is None:
            raise RuntimeError("warp_photometric_loss evaluates its gradients in the forward sweep and hands them out "
                               "once; for a second backward pass (retain_graph) use warp_photometric(...).mean()")
        grad_depth, grad_src, grad_P = ctx.grads
        ctx.grads = None
        K, T = ctx.saved_tensors
        need_depth, _, need_K, need_T, need_src = ctx.needs_input_grad[:5]
        g = f32(g, "grad").reshape(1).contiguous()
        with torch.cuda.device(grad_depth.device):
            rc = lib().e2e_scale_by_scalar(ptr(grad_depth), grad_depth.numel(),
                                           ptr(grad_src), grad_src.numel() if grad_src is not None else 0,
                                           ptr(grad_P), grad_P.numel() if grad_P is not None else 0,
                                           ptr(g), stream_ptr())
        check(rc, "e2e_scale_by_scalar")
        grad_K = grad_T = None
        if grad_P is not None:
            if need_T:
                grad_T = torch.matmul(K[:, :3, :].transpose(1, 2), grad_P)
            if need_K:
                grad_K = torch.zeros_like(K)
                grad_K[:, :3, :] = torch.matmul(grad_P, T.transpose(1, 2))
        return (grad_depth if need_depth else None, None, grad_K, grad_T, grad_src, None, None, None, None)


class _WarpPhotometricMeanDisp(torch.autograd.Function):
    """_WarpPhotometricMean fed with the depth network's disparity: depth = (1 / disp) * ratio is formed inside the sweep's depth
    load and d loss / d disp comes back directly (e2e_warp_photo_vg_disp) -- no depth tensor, no separate elementwise passes."""

    @staticmethod
    def forward(ctx, disp, ratio, inv_K, K, T, src, tgt, padding_mode, use_mask, eps):
        f32(disp, "disp"), f32(src, "source frame"), f32(tgt, "target frame")
        if disp.dim() != 4 or disp.shape[1] != 1:
            raise ValueError(f"disp must be (B,1,H,W), got {tuple(disp.shape)}")
        B, _, H, W = disp.shape
        if tuple(src.shape) != (B, 3, H, W) or tuple(tgt.shape) != (B, 3, H, W):
            raise ValueError(f"source/target frames must be ({B},3,{H},{W}), got {tuple(src.shape)} / {tuple(tgt.shape)}")
        if ratio is not None and ctx.needs_input_grad[1]:
            raise NotImplementedError("the median-scaling ratio is a constant of this op, like the reference's in-place `*= ratio`")
        disp_c = disp.contiguous()
        r = None if ratio is None else f32(ratio, "ratio").detach().reshape(1).contiguous()
        inv_K_c, K_c, T_c = _mat44(inv_K, B, "inv_K"), _mat44(K, B, "K"), _mat44(T, B, "T")
        prepare_divisors(W - 1, H - 1, 9.0, 3.0)
        dev = disp.device
        need_src, need_K, need_T = ctx.needs_input_grad[5], ctx.needs_input_grad[3], ctx.needs_input_grad[4]
        n = lib().e2e_warp_photo_vg_workspace_bytes(B, H, W)
        ws = torch.empty(n, dtype=torch.uint8, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        grad_disp = torch.empty_like(disp_c)
        grad_src = torch.zeros(B, 3, H, W, dtype=torch.float32, device=dev) if need_src else None
        grad_P = torch.empty(B, 3, 4, dtype=torch.float32, device=dev) if (need_K or need_T) else None
        with torch.cuda.device(dev):
            rc = lib().e2e_warp_photo_vg_disp(ptr(disp_c), ptr(r), ptr(inv_K_c), ptr(K_c), ptr(T_c), ptr(src), strides4(src),
                                              ptr(tgt), strides4(tgt), B, H, W, _pad_code(padding_mode), int(bool(use_mask)),
                                              ctypes.c_float(eps), ptr(loss), ptr(grad_disp), ptr(grad_src),
                                              strides4(grad_src) if grad_src is not None else None, ptr(grad_P), ptr(ws), n, stream_ptr())
        check(rc, "e2e_warp_photo_vg_disp")
        ctx.grads = (grad_disp, grad_src, grad_P)
        ctx.save_for_backward(K_c, T_c)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        if ctx.grads is None:
            raise RuntimeError("the gradients of this op are evaluated in the forward sweep and handed out once")
        grad_disp, grad_src, grad_P = ctx.grads
        ctx.grads = None
        K, T = ctx.saved_tensors
        g = f32(g, "grad").reshape(1).contiguous()
        with torch.cuda.device(grad_disp.device):
            check(lib().e2e_scale_by_scalar(ptr(grad_disp), grad_disp.numel(), ptr(grad_src), grad_src.numel() if grad_src is not None else 0,
                                            ptr(grad_P), grad_P.numel() if grad_P is not None else 0, ptr(g), stream_ptr()), "e2e_scale_by_scalar")
        grad_K = grad_T = None
        if grad_P is not None:
            if ctx.needs_input_grad[4]:
                grad_T = torch.matmul(K[:, :3, :].transpose(1, 2), grad_P)
            if ctx.needs_input_grad[3]:
                grad_K = torch.zeros_like(K)
                grad_K[:, :3, :] = torch.matmul(grad_P, T.transpose(1, 2))
        return (grad_disp if ctx.needs_input_grad[0] else None, None, None, grad_K, grad_T, grad_src, None, None, None, None)


def warp_photometric_loss_from_disparity(disp, inv_K, K, T, source_frame, target_frame, ratio=None, padding_mode="border",
                                         photometric_mask=True, eps=1e-7):
    """`warp_photometric_loss(1 / disp * ratio, ...)` with the conversion folded into the kernel (SURVEY.md 8(f) rank 2;
    online_adaption.py:282-298): `disp` (B,1,H,W) is the network's output, `ratio` the 0-dim device tensor of the median scaling
    (ops.median_ratio) or None.  Bit-identical loss, gradient w.r.t. the disparity returned directly."""
    return _WarpPhotometricMeanDisp.apply(disp, ratio, inv_K, K, T, source_frame, target_frame, padding_mode, photometric_mask, eps)


def warp_photometric(depth, inv_K, K, T, source_frame, target_frame, padding_mode="border",
                     photometric_mask=True, need_outputs=False, eps=1e-7, expect_uniform=True):
    """Per-pixel photometric loss of warping `source_frame` into the target view.

    Returns loss_map [B,1,H,W]; with need_outputs=True returns (loss_map, synthesized_frame [B,3,H,W],
    valid_mask [B,1,H,W], pixel_coordinates [B,H,W,2]) -- the tensors the reference keeps in `outputs`
    (train_depth.py:581-590).  Differentiable w.r.t. depth, K, T and source_frame.
    expect_uniform=False: the caller knows that the map will NOT be reduced by a plain mean (min-reprojection, auto-masking,
    per-pixel weights): the forward then skips the speculative gradients and backward runs the gradient-map sweep directly."""
    return _WarpPhotometric.apply(depth, inv_K, K, T, source_frame, target_frame, padding_mode,
                                  photometric_mask, eps, "map", need_outputs, bool(expect_uniform))


def warp_photometric_loss(depth, inv_K, K, T, source_frame, target_frame, padding_mode="border",
                          photometric_mask=True, eps=1e-7):
    """Scalar `photometric_loss(...).mean()` for one source frame (train_depth.py:657), lean path:
    neither the synthesized frame nor the loss map is written to memory."""
    return _WarpPhotometricMean.apply(depth, inv_K, K, T, source_frame, target_frame, padding_mode,
                                      photometric_mask, eps)


class _WarpPhotometricMeanMulti(torch.autograd.Function):
    """S source frames per target in ONE sweep launch (e2e_warp_photo_vg_multi): loss = mean over frames and pixels
    (`.mean(1, keepdim=True).mean()`, train_depth.py:629, 657) and all gradients, evaluated in forward()."""

    @staticmethod
    def forward(ctx, depth, inv_K, K, T, src, tgt, padding_mode, use_mask, eps):
        f32(depth, "depth"), f32(src, "source frames"), f32(tgt, "target frame")
        B, _, H, W = depth.shape
        if src.dim() != 5 or src.shape[0] != B or tuple(src.shape[2:]) != (3, H, W) or tuple(tgt.shape) != (B, 3, H, W):
            raise ValueError(f"expected sources ({B},S,3,{H},{W}) and target ({B},3,{H},{W}), got {tuple(src.shape)} / {tuple(tgt.shape)}")
        S = src.shape[1]
        if tuple(T.shape) != (B, S, 4, 4):
            raise ValueError(f"transforms must be ({B},{S},4,4), got {tuple(T.shape)}")
        if S > 1 and src.stride(0) != S * src.stride(1):          # (pair, source) must flatten to one strided axis
            src = src.contiguous()
        flat = src.flatten(0, 1) if S > 1 and src.stride(0) == S * src.stride(1) else src.reshape(B * S, 3, H, W)
        depth_c = depth.contiguous()
        inv_K_c, K_c = _mat44(inv_K, B, "inv_K"), _mat44(K, B, "K")
        T_c = f32(T, "T").reshape(B * S, 4, 4).contiguous()
        prepare_divisors(W - 1, H - 1, 9.0, 3.0)
        dev = depth.device
        need_depth, _, need_K, need_T, need_src = ctx.needs_input_grad[:5]
        n = lib().e2e_warp_photo_vg_multi_workspace_bytes(B, S, H, W)
        ws = torch.empty(n, dtype=torch.uint8, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        grad_depth = torch.empty_like(depth_c)
        grad_src = torch.zeros(B * S, 3, H, W, dtype=torch.float32, device=dev) if need_src else None
        grad_P = torch.empty(B * S, 3, 4, dtype=torch.float32, device=dev) if (need_K or need_T) else None
        with torch.cuda.device(dev):
            rc = lib().e2e_warp_photo_vg_multi(ptr(depth_c), ptr(inv_K_c), ptr(K_c), ptr(T_c), ptr(flat), strides4(flat), ptr(tgt),
                                               strides4(tgt), B, S, H, W, _pad_code(padding_mode), int(bool(use_mask)), ctypes.c_float(eps),
                                               ptr(loss), ptr(grad_depth), ptr(grad_src), strides4(grad_src) if grad_src is not None else None,
                                               ptr(grad_P), ptr(ws), n, stream_ptr())
        check(rc, "e2e_warp_photo_vg_multi")
        ctx.grads = (grad_depth, grad_src, grad_P)
        ctx.save_for_backward(K_c, T_c)
        ctx.dims = (B, S, H, W)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        if ctx.grads is None:
            raise RuntimeError("the gradients of this op are evaluated in the forward sweep and handed out once")
        grad_depth, grad_src, grad_P = ctx.grads
        ctx.grads = None
        K, T = ctx.saved_tensors
        B, S, H, W = ctx.dims
        need_depth, _, need_K, need_T, need_src = ctx.needs_input_grad[:5]
        g = f32(g, "grad").reshape(1).contiguous()
        with torch.cuda.device(grad_depth.device):
            check(lib().e2e_scale_by_scalar(ptr(grad_depth), grad_depth.numel(), ptr(grad_src), grad_src.numel() if grad_src is not None else 0,
                                            ptr(grad_P), grad_P.numel() if grad_P is not None else 0, ptr(g), stream_ptr()), "e2e_scale_by_scalar")
        grad_K = grad_T = None
        if grad_P is not None:
            Kr = K.repeat_interleave(S, 0)
            if need_T:
                grad_T = torch.matmul(Kr[:, :3, :].transpose(1, 2), grad_P).view(B, S, 4, 4)
            if need_K:
                grad_K = torch.zeros_like(K)
                grad_K[:, :3, :] = torch.matmul(grad_P, T.transpose(1, 2)).view(B, S, 3, 4).sum(1)
        return (grad_depth if need_depth else None, None, grad_K, grad_T,
                grad_src.view(B, S, 3, H, W) if grad_src is not None else None, None, None, None, None)


def warp_photometric_loss_multi(depth, inv_K, K, transforms, source_frames, target_frame, padding_mode="border", photometric_mask=True,
                                eps=1e-7):
    """Scalar photometric loss for S source frames per target -- `photometric_losses.mean(1, keepdim=True).mean()` of
    train_depth.py:726, 629, 657 -- in ONE launch of the sweep (grid z = pair x source).  `transforms` (B,S,4,4); `source_frames`
    (B,S,3,H,W), any inner strides (e.g. `colors[:, 1:].permute(0, 1, 4, 2, 3)` of channels-last memory goes in without a copy)."""
    return _WarpPhotometricMeanMulti.apply(depth, inv_K, K, transforms, source_frames, target_frame, padding_mode, photometric_mask, eps)


def warp_photometric_multi(depth, inv_K, K, transforms, source_frames, target_frame, padding_mode="border", photometric_mask=True,
                           min_reprojection=False, auto_masking=False, noise=None, eps=1e-7):
    """The reference's whole photometric objective for S source frames per target (novel_view_synthesis + compute_photometric_loss +
    compute_automasking_loss + the frame reduction of compute_losses; train_depth.py:545-613, 615-660, 707-750): one fused sweep per
    source frame (loss map + validity mask), the identity maps of the un-warped sources through the stand-alone SSIM kernel when
    `auto_masking`, and the mean / per-pixel-minimum reduction (losses.photometric_objective).  Returns the scalar loss;
    differentiable w.r.t. depth, the poses and the source frames.  With the plain mean (the shipped default) the upstream gradient
    of every map is uniform and the gradients come from the forward sweeps; under min-reprojection / auto-masking the per-pixel
    selection makes it non-uniform and each source frame's streaming backward kernel runs with the selection mask."""
    from . import losses
    if len(transforms) != len(source_frames) or not transforms:
        raise ValueError("one transform per source frame is required")
    maps, ident = [], []
    for T, src in zip(transforms, source_frames):
        if auto_masking:
            lm, _, valid, _ = warp_photometric(depth, inv_K, K, T, src, target_frame, padding_mode, photometric_mask, True, eps)
            if photometric_mask:                                  # train_depth.py:736-742
                ident.append(photometric_map(src * valid, target_frame * valid))
            else:
                ident.append(photometric_map(src, target_frame))
        else:
            lm = warp_photometric(depth, inv_K, K, T, src, target_frame, padding_mode, photometric_mask, False, eps)
        maps.append(lm)
    return losses.photometric_objective(maps, ident if auto_masking else None, min_reprojection, noise)


class _SSIM(torch.autograd.Function):
    """mode 'ssim' -> SSIM map [B,C,H,W] (losses.py:23-37);  mode 'photo' -> loss map [B,1,H,W] (:97-117)."""

    @staticmethod
    def forward(ctx, x, y, mode):
        f32(x, "x"), f32(y, "y")
        if x.shape != y.shape or x.dim() != 4:
            raise ValueError(f"x and y must be equal-shape (B,C,H,W) tensors, got {tuple(x.shape)} / {tuple(y.shape)}")
        B, C, H, W = x.shape
        prepare_divisors(9.0, 3.0, W - 1, H - 1)
        dev = x.device
        ssim_map = torch.empty(B, C, H, W, dtype=torch.float32, device=dev) if mode == "ssim" else None
        loss_map = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev) if mode == "photo" else None
        with torch.cuda.device(dev):
            rc = lib().e2e_ssim_fwd(ptr(x), strides4(x), ptr(y), strides4(y), B, C, H, W,
                                    ptr(ssim_map), ptr(loss_map), stream_ptr())
        check(rc, "e2e_ssim_fwd")
        ctx.save_for_backward(x, y)
        ctx.mode = mode
        return ssim_map if mode == "ssim" else loss_map

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        B, C, H, W = x.shape
        dev = x.device
        g = f32(g, "grad").contiguous()
        gx = torch.empty(B, C, H, W, dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        gy = torch.empty(B, C, H, W, dtype=torch.float32, device=dev) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(dev):
            rc = lib().e2e_ssim_bwd(ptr(x), strides4(x), ptr(y), strides4(y), B, C, H, W,
                                    ptr(g if ctx.mode == "ssim" else None), ptr(g if ctx.mode == "photo" else None),
                                    ptr(gx), ptr(gy), stream_ptr())
        check(rc, "e2e_ssim_bwd")
        return gx, gy, None


def ssim_map(x, y):
    return _SSIM.apply(x, y, "ssim")


def photometric_map(prediction, target):
    if prediction.dim() == 4 and prediction.shape[1] == 3:
        return _SSIM.apply(prediction, target, "photo")
    # other channel counts: SSIM kernel per plane, channel means are trivial torch reductions
    return 0.85 * ssim_map(prediction, target).mean(1, True) + 0.15 * torch.abs(target - prediction).mean(1, True)


def colors_from_uint8(frames_u8):
    """uint8 frames (any shape, contiguous, CUDA) -> float32 frames in [0, 1]: the reference's host-side `colors /= 255.0`
    (train_depth.py:255) done on the device, bit-identical to it, so that frames cross PCIe as bytes."""
    if frames_u8.dtype != torch.uint8:
        raise TypeError(f"expected a uint8 tensor, got {frames_u8.dtype}")
    if not frames_u8.is_cuda:
        raise _lib.E2ELibraryError("colors_from_uint8 needs a CUDA tensor (there is no CPU path)")
    src = frames_u8.contiguous()
    out = torch.empty(src.shape, dtype=torch.float32, device=src.device)
    if src.numel():
        with torch.cuda.device(src.device):
            check(lib().e2e_u8_to_unit(ptr(src), src.numel(), ptr(out), stream_ptr()), "e2e_u8_to_unit")
    return out


class _DispToDepth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, ratio):
        f32(disp, "disp")
        d = disp.contiguous()
        r = None if ratio is None else f32(ratio, "ratio").detach().reshape(1).contiguous()
        out = torch.empty_like(d)
        with torch.cuda.device(d.device):
            check(lib().e2e_disp_to_depth_fwd(ptr(d), ptr(r), d.numel(), ptr(out), stream_ptr()), "e2e_disp_to_depth_fwd")
        ctx.save_for_backward(d) if r is None else ctx.save_for_backward(d, r)
        return out

    @staticmethod
    def backward(ctx, g):
        saved = ctx.saved_tensors
        d, r = saved[0], (saved[1] if len(saved) > 1 else None)
        gd = torch.empty_like(d)
        with torch.cuda.device(d.device):
            check(lib().e2e_disp_to_depth_bwd(ptr(d), ptr(r), ptr(f32(g, "grad").contiguous()), d.numel(), ptr(gd), stream_ptr()),
                  "e2e_disp_to_depth_bwd")
        return gd, None


def disp_to_depth(disp, ratio=None):
    """depth = 1 / disp, optionally times the median-scaling ratio (a 0-dim device tensor, treated as a constant like the
    reference's in-place `*= ratio`): online_adaption.py:282, 295-298; train_depth.py:323-340.  One kernel each way."""
    return _DispToDepth.apply(disp, ratio)


def select_kth(x, k):
    """k-th smallest (0-based) element of a CUDA fp32 tensor as a 0-dim device tensor: radix select, no sort, no host sync."""
    f32(x, "x")
    flat = x.contiguous().view(-1)
    n = flat.numel()
    if not (0 <= k < n):
        raise IndexError(f"rank {k} out of range for {n} elements")
    out = torch.empty(1, dtype=torch.float32, device=x.device)
    nws = lib().e2e_select_workspace_bytes()
    ws = torch.empty(nws, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().e2e_select_kth(ptr(flat), n, k, ptr(out), ptr(ws), nws, stream_ptr()), "e2e_select_kth")
    return out.reshape(())


def median(x):
    """torch.median(x) over all elements (the lower median, NaN if any NaN) without torch's sort."""
    return select_kth(x, (x.numel() - 1) // 2)


def median_ratio_from_disparity(gt_depths, disps):
    """The reference's median scaling `ratio = torch.median(gt_depths) / torch.median(depth_tensor)` (online_adaption.py:295) with
    depth_tensor = 1 / disps never materialised: for positive disparities the lower median of 1/disp is exactly 1 / (the n/2-th
    smallest disparity).  Returns a 0-dim device tensor (a constant, like the reference's in-place `*= ratio`)."""
    with torch.no_grad():
        d = disps.detach()
        return median(gt_depths.detach()) / (1.0 / select_kth(d, d.numel() // 2))


def _row_mask(H, device):
    # l_mask of process_disparity (train_depth.py:231-234) along the rows; torch builds the H values, the kernel applies them
    return (1.0 - torch.clip(20 * (torch.linspace(0, 1, H) - 0.05), 0, 1)).to(device).contiguous()


class _DualDisparity(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp2):
        f32(disp2, "disp")
        if disp2.dim() != 4 or disp2.shape[0] != 2 or disp2.shape[1] != 1:
            raise ValueError(f"expected the disparities of a frame and of its mirror image, (2,1,H,W); got {tuple(disp2.shape)}")
        H, W = disp2.shape[2:]
        d = disp2.contiguous()
        m = _row_mask(H, d.device)
        out = torch.empty(1, 1, H, W, dtype=torch.float32, device=d.device)
        with torch.cuda.device(d.device):
            check(lib().e2e_dual_disparity_fwd(ptr(d[0]), ptr(d[1]), ptr(m), H, W, ptr(out), stream_ptr()), "e2e_dual_disparity_fwd")
        ctx.save_for_backward(m)
        return out

    @staticmethod
    def backward(ctx, g):
        (m,) = ctx.saved_tensors
        H, W = g.shape[2:]
        gd = torch.empty(2, 1, H, W, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            check(lib().e2e_dual_disparity_bwd(ptr(f32(g, "grad").contiguous()), ptr(m), H, W, ptr(gd[0]), ptr(gd[1]), stream_ptr()),
                  "e2e_dual_disparity_bwd")
        return gd


def dual_disparity(disp2):
    """process_disparity (train_depth.py:224-237): disp2 (2,1,H,W) = the depth network's disparities for a frame and for its
    mirror image (the reference feeds `torch.cat([frame, torch.flip(frame, [2])], 0)`, train_depth.py:322-327); the second
    one is flipped back along W and blended with the first under the row mask.  Returns (1,1,H,W); differentiable."""
    return _DualDisparity.apply(disp2)


def launch_count():
    return _lib.launch_count()


class WarpPhotoPlan:
    """Pre-allocated, autograd-free driver of the fused kernels for a fixed (B, H, W): the form a training
    loop (or a CUDA graph) wants -- no allocation, no host synchronisation, two kernel launches forward
    (fused kernel + partial-sum reduce) and two backward (fused kernel + grad_P reduce).

        plan = WarpPhotoPlan(B, H, W, device)
        loss = plan.forward(depth, inv_K, K, T, src, tgt)              # device scalar: mean photometric loss
        g_depth, g_src, g_P = plan.backward(depth, inv_K, K, T, src, tgt)   # d loss / d {depth, src, (K@T)[:3]}
    """

    def __init__(self, B, H, W, device, padding_mode="border", photometric_mask=True, eps=1e-7,
                 need_src_grad=True, need_pose_grad=True, overlap_zero_fill=False):
        self.B, self.H, self.W = B, H, W
        self.device = torch.device(device)
        self.pad, self.mask, self.eps = _pad_code(padding_mode), int(bool(photometric_mask)), float(eps)
        with torch.cuda.device(self.device):
            prepare_divisors(W - 1, H - 1, 9.0, 3.0)
        self.ws, self.ws_bytes = _workspace(B, H, W, self.device)
        self.vg_ws_bytes = lib().e2e_warp_photo_vg_workspace_bytes(B, H, W)
        self.vg_ws = torch.empty(self.vg_ws_bytes, dtype=torch.uint8, device=self.device)
        f = dict(dtype=torch.float32, device=self.device)
        self.loss = torch.empty(1, **f)
        self.grad_depth = torch.empty(B, 1, H, W, **f)
        # planar: the kernel's atomics of a warp then fall into 4 sectors per instruction instead of 12
        self.grad_src = torch.empty(B, 3, H, W, **f) if need_src_grad else None
        self.grad_P = torch.empty(B, 3, 4, **f) if need_pose_grad else None
        self._gs_strides = strides4(self.grad_src) if need_src_grad else None
        # overlap_zero_fill: two grad_src buffers; while the kernel of this call accumulates into one, the other is cleared on a
        # side stream for the next call (the kernel is issue-bound, the memset costs it nothing).  The buffer returned by a call
        # stays valid until the next call starts, as without the option.  Not for CUDA-graph capture (the side stream is not joined).
        self._overlap = bool(overlap_zero_fill and need_src_grad)
        if self._overlap:
            with torch.cuda.device(self.device):
                self._bufs = [self.grad_src, torch.empty(B, 3, H, W, **f)]
                self._side = torch.cuda.Stream(device=self.device)
                self._zero_done = [torch.cuda.Event(), torch.cuda.Event()]
                self._cur = 0
                main = torch.cuda.current_stream(self.device)
                for b_, e_ in zip(self._bufs, self._zero_done):
                    b_.zero_()
                    e_.record(main)

    def value_and_grad(self, depth, inv_K, K, T, src, tgt):
        """loss (device scalar) and d loss / d {depth, src, (K@T)[:3]} in ONE sweep (e2e_warp_photo_vg):
        zero-fill of grad_src + fused kernel + two tiny fixed-order reductions."""
        if self._overlap:
            main = torch.cuda.current_stream(self.device)
            i = self._cur
            self.grad_src = self._bufs[i]
            main.wait_event(self._zero_done[i])
            start = torch.cuda.Event()
            start.record(main)                       # everything that still reads the other buffer was enqueued before this point
            with torch.cuda.stream(self._side):
                self._side.wait_event(start)
                self._bufs[1 - i].zero_()
                self._zero_done[1 - i].record(self._side)
            self._cur = 1 - i
        elif self.grad_src is not None:
            self.grad_src.zero_()          # the kernel accumulates into it with red.global.add
        with torch.cuda.device(self.device):
            rc = lib().e2e_warp_photo_vg(ptr(depth), ptr(inv_K), ptr(K), ptr(T), ptr(src), strides4(src),
                                         ptr(tgt), strides4(tgt), self.B, self.H, self.W, self.pad, self.mask,
                                         ctypes.c_float(self.eps), ptr(self.loss), ptr(self.grad_depth),
                                         ptr(self.grad_src), self._gs_strides, ptr(self.grad_P),
                                         ptr(self.vg_ws), self.vg_ws_bytes, stream_ptr())
        check(rc, "e2e_warp_photo_vg")
        return self.loss, self.grad_depth, self.grad_src, self.grad_P

    def forward(self, depth, inv_K, K, T, src, tgt):
        with torch.cuda.device(self.device):
            rc = lib().e2e_warp_photo_fwd(ptr(depth), ptr(inv_K), ptr(K), ptr(T), ptr(src), strides4(src),
                                          ptr(tgt), strides4(tgt), self.B, self.H, self.W, self.pad, self.mask,
                                          ctypes.c_float(self.eps), None, None, None, None, ptr(self.loss),
                                          ptr(self.ws), self.ws_bytes, stream_ptr())
        check(rc, "e2e_warp_photo_fwd")
        return self.loss

    def backward(self, depth, inv_K, K, T, src, tgt, grad_loss=None):
        if self.grad_src is not None:
            self.grad_src.zero_()          # the kernel accumulates into it with red.global.add
        with torch.cuda.device(self.device):
            rc = lib().e2e_warp_photo_bwd(ptr(depth), ptr(inv_K), ptr(K), ptr(T), ptr(src), strides4(src),
                                          ptr(tgt), strides4(tgt), self.B, self.H, self.W, self.pad, self.mask,
                                          ctypes.c_float(self.eps), None, ptr(grad_loss),
                                          ctypes.c_float(1.0 / (self.B * self.H * self.W)),
                                          ptr(self.grad_depth), ptr(self.grad_src), self._gs_strides,
                                          ptr(self.grad_P), ptr(self.ws), self.ws_bytes, stream_ptr())
        check(rc, "e2e_warp_photo_bwd")
        return self.grad_depth, self.grad_src, self.grad_P
