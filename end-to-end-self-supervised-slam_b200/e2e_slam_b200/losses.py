"""Drop-in for the reference's loss/losses.py: same function names, arguments, return values and error
behaviour, backed by the CUDA kernels (no chamferdist / matplotlib imports needed).

    SSIM, photometric_loss, disparity_smoothness_loss, depth_reguralizer (sic), depth_gt_loss,
    geometric_consistency_loss, knn_points_loss, color_points_loss, depth_metrics, compute_depth_errors
plus `smoothness_loss(disp, img)`, the fused form of compute_smoothness_loss (train_depth.py:763-773).
"""
import ctypes
import os

import torch
import torch.nn as nn

from . import ops
from ._lib import check, f32, lib, ptr, stream_ptr, strides4


def _red_ws(device):
    n = lib().e2e_reduce_workspace_bytes(0)
    return torch.empty(n, dtype=torch.uint8, device=device), n


class SSIM(nn.Module):
    """SSIM loss map between two images (losses.py:6-37): reflect-pad 1, 3x3 mean pools,
    clamp((1 - SSIM)/2, 0, 1).  forward(x, y) -> [B,C,H,W]."""

    def __init__(self):
        super().__init__()
        self.C1 = 0.01 ** 2
        self.C2 = 0.03 ** 2

    def forward(self, x, y):
        return ops.ssim_map(x, y)


def photometric_loss(ssim, prediction, target):
    """0.85 * mean_c SSIM + 0.15 * mean_c |target - prediction|  -> [B,1,H,W]   (losses.py:97-117).
    With an e2e_slam_b200 SSIM instance the whole expression is one kernel; any other callable `ssim`
    is honoured as given (the reference passes the module in)."""
    if isinstance(ssim, SSIM):
        return ops.photometric_map(prediction, target)
    ssim_loss = ssim(x=prediction, y=target).mean(1, True)
    return 0.85 * ssim_loss + 0.15 * torch.abs(target - prediction).mean(1, True)


_SMOOTH_FUSED = os.environ.get("E2E_SMOOTH_FUSED", "1") != "0"


class _Smooth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, img, raw=False):
        """raw: the disparity is used as given (disparity_smoothness_loss of an already normalised disparity)."""
        f32(disp, "disp"), f32(img, "img")
        if disp.dim() != 4 or disp.shape[1] != 1 or img.dim() != 4 or img.shape[1] != 3 or img.shape[2:] != disp.shape[2:]:
            raise ValueError(f"expected disp (B,1,H,W) and img (B,3,H,W), got {tuple(disp.shape)} / {tuple(img.shape)}")
        B, _, H, W = disp.shape
        disp_c = disp.contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=disp.device)
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("the smoothness loss is not differentiated w.r.t. the image (the reference never does)")
        ctx.fused = (bool(ctx.needs_input_grad[0]) and _SMOOTH_FUSED and H >= 2 and W >= 2) or raw
        if ctx.fused:        # value and d loss / d n in one sweep; backward is one elementwise pass
            gn = torch.empty_like(disp_c)
            stats = torch.empty(B, 2, dtype=torch.float32, device=disp.device)
            n = lib().e2e_smooth_vg_workspace_bytes(B, H, W)
            ws = torch.empty(n, dtype=torch.uint8, device=disp.device)
            with torch.cuda.device(disp.device):
                fn = lib().e2e_smooth_vg_raw if raw else lib().e2e_smooth_vg
                check(fn(ptr(disp_c), ptr(img), strides4(img), B, H, W, ptr(loss), ptr(gn), ptr(stats), ptr(ws), n, stream_ptr()), "e2e_smooth_vg")
            ctx.save_for_backward(gn, stats)
            return loss.reshape(())
        ws, n = _red_ws(disp.device)
        with torch.cuda.device(disp.device):
            check(lib().e2e_smooth_fwd(ptr(disp_c), ptr(img), strides4(img), B, H, W, ptr(loss), ptr(ws), n, stream_ptr()),
                  "e2e_smooth_fwd")
        ctx.save_for_backward(disp_c, img)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        if ctx.fused:
            gn, stats = ctx.saved_tensors
            B, _, H, W = gn.shape
            gd = torch.empty_like(gn)
            g = f32(g, "grad").reshape(1).contiguous()
            with torch.cuda.device(gn.device):
                check(lib().e2e_smooth_apply(ptr(gn), ptr(stats), ptr(g), B, H, W, ptr(gd), stream_ptr()), "e2e_smooth_apply")
            return gd, None, None
        disp, img = ctx.saved_tensors
        B, _, H, W = disp.shape
        gd = torch.empty_like(disp)
        ws, n = _red_ws(disp.device)
        g = f32(g, "grad").reshape(1).contiguous()
        with torch.cuda.device(disp.device):
            check(lib().e2e_smooth_bwd(ptr(disp), ptr(img), strides4(img), B, H, W, ptr(g), ptr(gd), ptr(ws), n, stream_ptr()),
                  "e2e_smooth_bwd")
        return gd, None, None


def smoothness_loss(disp, img):
    """compute_smoothness_loss (train_depth.py:763-773): mean-normalise the disparity per image, then the
    edge-aware smoothness of losses.py:119-132, fused.  Differentiable w.r.t. disp."""
    return _Smooth.apply(disp, img)


def disparity_smoothness_loss(disp, img):
    """Edge-aware smoothness of an ALREADY normalised disparity (losses.py:119-132), same signature as the reference: what the
    unmodified compute_smoothness_loss (train_depth.py:763-773) calls after its own normalisation.  One sweep (value and
    d loss / d disp, e2e_smooth_vg_raw) + one elementwise pass in backward; scripts that can should call `smoothness_loss`,
    which also fuses the normalisation."""
    return _Smooth.apply(disp, img, True)


class _EwLoss(torch.autograd.Function):
    """kind: 'sparse' (pred, mask, gt) | 'l1' / 'l2' (initial, refined)."""

    @staticmethod
    def forward(ctx, kind, a, b, c):
        n = a.numel()
        loss = torch.empty(1, dtype=torch.float32, device=a.device)
        ws, nb = _red_ws(a.device)
        with torch.cuda.device(a.device):
            if kind == "sparse":
                rc = lib().e2e_sparse_l1_fwd(ptr(a), ptr(b), ptr(c), n, ptr(loss), ptr(ws), nb, stream_ptr())
            else:
                rc = lib().e2e_depth_reg_fwd(ptr(a), ptr(b), n, 1 if kind == "l1" else 2, ptr(loss), ptr(ws), nb, stream_ptr())
        check(rc, "elementwise loss")
        ctx.kind = kind
        ctx.save_for_backward(a, b) if c is None else ctx.save_for_backward(a, b, c)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        kind = ctx.kind
        g = f32(g, "grad").reshape(1).contiguous()
        if kind == "sparse":
            a, b, c = ctx.saved_tensors
            ga = torch.empty_like(a)
            with torch.cuda.device(a.device):
                check(lib().e2e_sparse_l1_bwd(ptr(a), ptr(b), ptr(c), a.numel(), ptr(g), ptr(ga), stream_ptr()), "e2e_sparse_l1_bwd")
            return None, ga, None, None
        a, b = ctx.saved_tensors
        gb = torch.empty_like(b)
        with torch.cuda.device(a.device):
            check(lib().e2e_depth_reg_bwd(ptr(a), ptr(b), a.numel(), 1 if kind == "l1" else 2, ptr(g), ptr(gb), stream_ptr()),
                  "e2e_depth_reg_bwd")
        return None, None, gb, None


def depth_reguralizer(initial_depth, refined_depth, loss_func):
    """mean |initial - refined| ('l1') or mean (initial - refined)^2 ('l2')   (losses.py:134-148).
    Gradient flows to refined_depth (initial_depth is a detached clone in the reference, train_depth.py:336)."""
    if loss_func not in ("l1", "l2"):
        raise ValueError("please specify a correct norm")
    f32(initial_depth, "initial_depth"), f32(refined_depth, "refined_depth")
    if initial_depth.shape != refined_depth.shape:
        raise ValueError("initial and refined depth must have the same shape")
    if torch.is_grad_enabled() and initial_depth.requires_grad:
        raise NotImplementedError("depth_reguralizer differentiates refined_depth only (the reference passes a detached clone as "
                                  "initial_depth, train_depth.py:336); detach it or swap the arguments")
    return _EwLoss.apply(loss_func, initial_depth.detach().contiguous(), refined_depth.contiguous(), None)


def depth_gt_loss(prediction, sparse_groundtruth, sparse_mask):
    """L1 between prediction*mask and the sparse ground truth, averaged over ALL pixels (losses.py:151-160).
    The reference `.squeeze()`s all three tensors; as there, the element counts must agree."""
    f32(prediction, "prediction"), f32(sparse_groundtruth, "sparse_groundtruth"), f32(sparse_mask, "sparse_mask")
    p, g, m = prediction.squeeze(), sparse_groundtruth.squeeze(), sparse_mask.squeeze()
    if p.shape != g.shape or p.shape != m.shape:
        raise ValueError(f"shapes after squeeze() differ: {tuple(p.shape)}, {tuple(g.shape)}, {tuple(m.shape)}")
    return _EwLoss.apply("sparse", p.contiguous(), m.detach().contiguous(), g.detach().contiguous())


class _Geometric(torch.autograd.Function):
    @staticmethod
    def forward(ctx, wd, idp, mask):
        wd_c, id_c, m_c = wd.contiguous(), idp.contiguous(), mask.contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=wd.device)
        ws, nb = _red_ws(wd.device)
        with torch.cuda.device(wd.device):
            check(lib().e2e_geometric_fwd(ptr(wd_c), ptr(id_c), ptr(m_c), wd_c.numel(), ptr(loss), ptr(ws), nb, stream_ptr()),
                  "e2e_geometric_fwd")
        ctx.save_for_backward(wd_c, id_c, m_c)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        wd, idp, m = ctx.saved_tensors
        g = f32(g, "grad").reshape(1).contiguous()
        msum = m.sum().reshape(1)                       # stays on the device: the > 10000 test is evaluated by the kernel
        ga = torch.empty_like(wd) if ctx.needs_input_grad[0] else None
        gb = torch.empty_like(idp) if ctx.needs_input_grad[1] else None
        if ga is None and gb is None:
            return None, None, None
        with torch.cuda.device(wd.device):
            check(lib().e2e_geometric_bwd(ptr(wd), ptr(idp), ptr(m), wd.numel(), ptr(msum), ptr(g), ptr(ga), ptr(gb), stream_ptr()),
                  "e2e_geometric_bwd")
        return ga, gb, None


def geometric_consistency_loss(outputs, frame, device):
    """clamp(|wd - id| / (wd + id), 0, 1) averaged over the valid mask if it has > 10000 pixels, else 0
    (losses.py:84-95), differentiable w.r.t. both depth maps.  The count test runs on the device (the reference syncs)."""
    wd, idp = outputs[("warped_depth", frame)], outputs[("interpolated_depth", frame)]
    mask = outputs[("valid_mask", frame)].expand_as(wd)
    f32(wd, "warped_depth"), f32(idp, "interpolated_depth"), f32(mask, "valid_mask")
    return _Geometric.apply(wd, idp, mask.detach())


@torch.no_grad()
def compute_depth_errors(gt, pred):
    """Evaluation metrics (losses.py:183-201): trivial reductions, kept as torch ops on the GPU."""
    thresh = torch.max((gt / pred), (pred / gt))
    a1 = (thresh < 1.25).float().mean()
    a2 = (thresh < 1.25 ** 2).float().mean()
    a3 = (thresh < 1.25 ** 3).float().mean()
    rmse = torch.sqrt(((gt - pred) ** 2).mean())
    rmse_log = torch.sqrt(((torch.log(gt) - torch.log(pred)) ** 2).mean())
    abs_rel = torch.mean(torch.abs(gt - pred) / gt)
    sq_rel = torch.mean((gt - pred) ** 2 / gt)
    return abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3


@torch.no_grad()
def depth_metrics(dataset, gt, pred):
    """losses.py:162-181."""
    pred, gt = pred.squeeze().detach(), gt.squeeze().detach()
    if dataset == "TUM":
        valid = gt != 0.0
    elif dataset == "ICL":
        valid = torch.ones_like(gt, dtype=torch.bool)
    else:
        raise ValueError("Dataset Not Found")
    return compute_depth_errors(gt[valid], pred[valid])


# ------------------------------------------------------------------------------------------------
# Multi-source photometric objective: mean over frames, min-reprojection, auto-masking (train_depth.py:615-660)
# ------------------------------------------------------------------------------------------------
class _MinComposite(torch.autograd.Function):
    """mean over pixels of the per-pixel minimum over the candidate maps (`torch.min(photmetric, dim=1)` then `.mean()`,
    train_depth.py:657-658), without materialising the reference's torch.cat; backward routes 1/n to the winning candidate."""

    @staticmethod
    def forward(ctx, *cands):
        c0 = cands[0]
        maps = [f32(c, "candidate map").contiguous() for c in cands]
        if any(m.shape != c0.shape for m in maps):
            raise ValueError("candidate maps must have equal shapes")
        if not (1 <= len(maps) <= 8):
            raise ValueError("between 1 and 8 candidate maps are supported")
        n = c0.numel()
        dev = c0.device
        index = torch.empty(n, dtype=torch.uint8, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        nws = lib().e2e_reduce_workspace_bytes(n)
        ws = torch.empty(nws, dtype=torch.uint8, device=dev)
        arr = (ctypes.c_void_p * len(maps))(*[m.data_ptr() for m in maps])
        with torch.cuda.device(dev):
            check(lib().e2e_min_composite_fwd(arr, len(maps), n, ptr(index), ptr(loss), ptr(ws), nws, stream_ptr()), "e2e_min_composite_fwd")
        ctx.save_for_backward(index)
        ctx.shape, ctx.k = c0.shape, len(maps)
        ctx.mark_non_differentiable(index)
        return loss.reshape(()), index.view(c0.shape)

    @staticmethod
    def backward(ctx, g, _):
        (index,) = ctx.saved_tensors
        g = f32(g, "grad").reshape(1).contiguous()
        grads = [torch.empty(ctx.shape, dtype=torch.float32, device=index.device) if ctx.needs_input_grad[k] else None for k in range(ctx.k)]
        arr = (ctypes.c_void_p * ctx.k)(*[0 if t is None else t.data_ptr() for t in grads])
        with torch.cuda.device(index.device):
            check(lib().e2e_min_composite_bwd(ptr(index), ctx.k, index.numel(), ptr(g), arr, stream_ptr()), "e2e_min_composite_bwd")
        return tuple(grads)


def photometric_objective(photometric_maps, identity_maps=None, min_reprojection=False, noise=None, return_index=False):
    """The scalar the reference optimises from the per-source-frame photometric maps (compute_losses, train_depth.py:621-660):
        photometric_maps   list of S maps (B,1,H,W) -- compute_photometric_loss, :707-727
        identity_maps      list of S maps of the UN-warped source (compute_automasking_loss, :729-750) when LOSS.auto_masking
        min_reprojection   LOSS.min_reprojection: keep the S maps as separate candidates instead of averaging them (:624-629)
        noise              (B,S,H,W) tie-breaking noise added to the identity maps under min_reprojection (:646); the reference draws
                           torch.randn(...) * 1e-5 on the spot -- pass it in for reproducibility (default: drawn here)
    One candidate -> its mean; several -> mean of the per-pixel minimum (one kernel, e2e_min_composite_fwd)."""
    maps = list(photometric_maps)
    if not maps:
        raise ValueError("no photometric maps")
    cands = maps if min_reprojection else [maps[0] if len(maps) == 1 else torch.cat(maps, 1).mean(1, keepdim=True)]
    if identity_maps is not None:
        ident = list(identity_maps)
        if len(ident) != len(maps):
            raise ValueError("one identity map per source frame is required")
        if min_reprojection:
            if noise is None:
                noise = torch.randn(maps[0].shape[0], len(ident), *maps[0].shape[2:], device=maps[0].device) * 0.00001
            ident = [m + noise[:, s:s + 1] for s, m in enumerate(ident)]
        else:
            ident = [ident[0] if len(ident) == 1 else torch.cat(ident, 1).mean(1, keepdim=True)]
        cands = ident + cands                                  # torch.cat((auto_masking, photmetric), dim=1), :651
    if len(cands) == 1:
        loss, index = cands[0].mean(), None
    else:
        loss, index = _MinComposite.apply(*cands)
    return (loss, index) if return_index else loss


# ------------------------------------------------------------------------------------------------
# Point supervision: chamferdist.chamfer.knn_points (K = 1) and the losses built on it
# ------------------------------------------------------------------------------------------------
from collections import namedtuple  # noqa: E402

_KNN = namedtuple("KNN", "dists idx knn")


KNN_MODE = os.environ.get("E2E_KNN", "auto")        # "auto" | "brute" | "grid" (tests and A/B timing)


def _use_grid(P1, P2):
    if P1 > (1 << 24) or P2 >= (1 << 31):           # limits of the grid entry points (far-query list, 32-bit point indices)
        return False
    if KNN_MODE == "auto":
        return P1 * P2 >= (1 << 24)                  # measured: 19 200 x 75 000 -> grid 0.15 ms, brute force 5.7 ms
    return KNN_MODE == "grid"


_GRID_CACHE = {}


def _grid_for(ref):
    """The grid over a reference cloud, rebuilt only when the cloud changes: the global map is the same tensor for every
    refinement step on a key-frame (online_adaption.py:638-645 detaches it).  The entry keeps the tensor alive, so its
    storage address cannot be recycled, and any in-place write bumps the version counter that is part of the key."""
    key = (ref.untyped_storage().data_ptr(), ref.storage_offset(), tuple(ref.shape), tuple(ref.stride()), ref._version, ref.device)
    hit = _GRID_CACHE.get("last")
    if hit is not None and hit[0] == key:
        return hit[2]
    _GRID_CACHE.pop("last", None)                     # free the old grid before allocating the new one
    P2 = ref.shape[0]
    nws = lib().e2e_knn1_grid_workspace_bytes(P2)
    ws = torch.empty(nws, dtype=torch.uint8, device=ref.device)
    check(lib().e2e_knn1_grid_build(ptr(ref), P2, ptr(ws), nws, stream_ptr()), "e2e_knn1_grid_build")
    _GRID_CACHE["last"] = (key, ref, ws)
    return ws


class _KNN1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, query, ref, transform):
        f32(query, "query points"), f32(ref, "reference points")
        q, r = query.contiguous(), ref.contiguous()
        t = None if transform is None else f32(transform, "transform").contiguous()
        P1, P2 = q.shape[0], r.shape[0]
        if ctx.needs_input_grad[2]:
            raise NotImplementedError("nearest-neighbour distances are not differentiated w.r.t. the fused transform "
                                      "(the reference detaches the pose, online_adaption.py:640-642); use slam.transform_pointcloud")
        if P2 == 0:
            raise ValueError("knn_points: the reference cloud is empty")
        dist2 = torch.empty(P1, dtype=torch.float32, device=q.device)
        idx = torch.empty(P1, dtype=torch.int64, device=q.device)
        if P1 == 0:                                       # e.g. an all-invalid live depth map: nothing to match
            ctx.save_for_backward(q, r, idx) if t is None else ctx.save_for_backward(q, r, idx, t)
            ctx.mark_non_differentiable(idx)
            return dist2, idx
        with torch.cuda.device(q.device):
            if _use_grid(P1, P2):      # same answer bit for bit, cost ~ P1 + P2 instead of P1 * P2
                ws = _grid_for(r)
                check(lib().e2e_knn1_grid_query(ptr(q), ptr(t), P1, P2, ptr(dist2), ptr(idx), ptr(ws), stream_ptr()),
                      "e2e_knn1_grid_query")
            else:
                check(lib().e2e_knn1_fwd(ptr(q), ptr(t), ptr(r), P1, P2, ptr(dist2), ptr(idx), stream_ptr()), "e2e_knn1_fwd")
        ctx.save_for_backward(q, r, idx) if t is None else ctx.save_for_backward(q, r, idx, t)
        ctx.mark_non_differentiable(idx)
        return dist2, idx

    @staticmethod
    def backward(ctx, g, _):
        saved = ctx.saved_tensors
        q, r, idx = saved[:3]
        t = saved[3] if len(saved) > 3 else None
        g = f32(g, "grad").contiguous()
        gq = torch.empty_like(q) if ctx.needs_input_grad[0] else None
        gr = torch.zeros_like(r) if ctx.needs_input_grad[1] else None
        if (gq is None and gr is None) or q.shape[0] == 0:
            return gq, gr, None
        with torch.cuda.device(q.device):
            check(lib().e2e_knn1_bwd(ptr(q), ptr(t), ptr(r), q.shape[0], r.shape[0], ptr(idx), ptr(g), ptr(gq), ptr(gr), stream_ptr()),
                  "e2e_knn1_bwd")
        return gq, gr, None


def knn_points(p1, p2, lengths1=None, lengths2=None, K=1, version=-1, return_nn=False, return_sorted=True):
    """chamferdist.chamfer.knn_points for K = 1: for every point of p1 (N,P1,3) the squared distance to and
    the index of its nearest neighbour in p2 (N,P2,3).  Returns KNN(dists (N,P1,1), idx (N,P1,1), knn=None)."""
    if K != 1 or lengths1 is not None or lengths2 is not None or return_nn:
        raise NotImplementedError("only K = 1 on full clouds is used by the reference (loss/losses.py:57)")
    if p1.dim() != 3 or p2.dim() != 3 or p1.shape[2] != 3 or p2.shape[2] != 3:
        raise ValueError(f"expected (N,P,3) point clouds, got {tuple(p1.shape)} / {tuple(p2.shape)}")
    if p1.shape[0] != p2.shape[0]:
        raise ValueError("pts1 and pts2 must have the same batch dimension.")
    d, i = zip(*[_KNN1.apply(p1[b], p2[b], None) for b in range(p1.shape[0])])
    return _KNN(dists=torch.stack(d).unsqueeze(-1), idx=torch.stack(i).unsqueeze(-1), knn=None)


def knn_points_loss(gt_pointcloud, noisy_pointcloud):
    """Mean squared distance from every noisy point to its nearest ground-truth point, and the indices
    (loss/losses.py:39-63).  Same argument order and errors as the reference."""
    if gt_pointcloud.shape[0] != noisy_pointcloud.shape[0]:
        raise ValueError("Pointclouds must have the same batch dimension")
    if gt_pointcloud.shape[2] != noisy_pointcloud.shape[2]:
        raise ValueError("Number of axes is not the same in both pointclouds")
    KNN = knn_points(noisy_pointcloud, gt_pointcloud)
    distances = KNN.dists.squeeze(-1)
    indexes = KNN.idx.squeeze(-1).detach()
    return torch.mean(distances), indexes


class _ColorPoints(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gt_col, noisy_col, idx):
        gt, noisy, idx = gt_col.contiguous(), noisy_col.contiguous(), idx.contiguous()
        P1 = noisy.shape[0]
        loss = torch.empty(1, dtype=torch.float32, device=noisy.device)
        nws = lib().e2e_color_points_workspace_bytes(P1)
        ws = torch.empty(nws, dtype=torch.uint8, device=noisy.device)
        with torch.cuda.device(noisy.device):
            check(lib().e2e_color_points_fwd(ptr(gt), ptr(noisy), ptr(idx), P1, ptr(loss), ptr(ws), nws, stream_ptr()), "e2e_color_points_fwd")
        ctx.save_for_backward(gt, noisy, idx)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        gt, noisy, idx = ctx.saved_tensors
        g = f32(g, "grad").reshape(1).contiguous()
        g_noisy = torch.empty_like(noisy) if ctx.needs_input_grad[1] else None
        g_gt = torch.zeros_like(gt) if ctx.needs_input_grad[0] else None
        if g_noisy is None and g_gt is None:
            return None, None, None
        with torch.cuda.device(noisy.device):
            check(lib().e2e_color_points_bwd(ptr(gt), ptr(noisy), ptr(idx), noisy.shape[0], ptr(g), ptr(g_noisy), ptr(g_gt), stream_ptr()),
                  "e2e_color_points_bwd")
        return g_gt, g_noisy, None


def color_points_loss(gt_pointcloud_color, noisy_pointcloud_color, indexes):
    """L1 between each noisy point's colour and the colour of its matched ground-truth point
    (loss/losses.py:65-82): gather + |.| + mean in one kernel (e2e_color_points_fwd), one kernel back."""
    if gt_pointcloud_color.shape[2] != noisy_pointcloud_color.shape[2]:
        raise ValueError("Number of axes is not the same in both pointclouds")
    f32(gt_pointcloud_color, "gt_pointcloud_color"), f32(noisy_pointcloud_color, "noisy_pointcloud_color")
    if gt_pointcloud_color.shape[2] != 3:
        return torch.mean(torch.abs(noisy_pointcloud_color[0] - gt_pointcloud_color[0, indexes[0].long()]))
    idx = indexes[0].long()
    if idx.shape[0] != noisy_pointcloud_color.shape[1]:
        raise ValueError(f"indexes ({idx.shape[0]}) must have one entry per noisy point ({noisy_pointcloud_color.shape[1]})")
    if idx.shape[0] == 0:
        return torch.full((), float("nan"), dtype=torch.float32, device=noisy_pointcloud_color.device)      # mean of an empty tensor
    return _ColorPoints.apply(gt_pointcloud_color[0], noisy_pointcloud_color[0], idx)


def point_supervision_loss(target_points, transform, global_points):
    """compute_3d_loss (online_adaption.py:638-645) with transform_pointcloud fused into the kernel's query
    load: mean squared distance of R p + t (p in target_points (N,3)) to the nearest of global_points (M,3),
    which is treated as constant like the reference's `.detach()`."""
    d, _ = _KNN1.apply(target_points, global_points.detach(), transform)
    return d.mean()


class ChamferDistance(nn.Module):
    """chamferdist.ChamferDistance, forward direction plus optional reverse (train_depth.py:689-695)."""

    def forward(self, source_cloud, target_cloud, bidirectional=False, reverse=False, reduction="mean"):
        if reduction not in ("mean", "sum"):
            raise ValueError('reduction must be "mean" or "sum"')
        red = (lambda t: t.mean(1).mean()) if reduction == "mean" else (lambda t: t.sum())
        fwd = red(knn_points(source_cloud, target_cloud).dists.squeeze(-1))
        if not (bidirectional or reverse):
            return fwd
        bwd = red(knn_points(target_cloud, source_cloud).dists.squeeze(-1))
        return fwd + bwd if bidirectional else bwd
