"""CPU oracle for the warp + photometric-loss path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module.  The product package (`e2e_slam_b200`) never does, and it has
no CPU fallback: without its CUDA library it raises.

This is a restatement of the reference's algorithm in plain torch CPU ops, op for op in the
reference's order, so that on CPU it reproduces the reference bit for bit (pinned by
tests/test_oracle_golden.py against tests/golden/*.npz, which were produced by importing the
real reference -- see tools/make_golden.py).  Gradients come from torch.autograd applied to
this forward, i.e. they are NOT hand-derived and therefore independent of the CUDA backward.

Reference lines restated (paths relative to the reference repo root):
  backproject            depth_estimation/view_synthesis.py:34-40  (pixel grid :17-31)
  project                depth_estimation/view_synthesis.py:54-78
  sample                 train_depth.py:587-590  (F.grid_sample, align_corners=False)
  mask multiply          train_depth.py:713-718
  ssim                   loss/losses.py:23-37
  photometric            loss/losses.py:97-117
  frames mean / min      train_depth.py:629, 657-660
  min-reprojection, auto-masking objective   train_depth.py:615-660, 729-750
  smoothness             train_depth.py:763-773 + loss/losses.py:119-132
  sparse gt L1           loss/losses.py:151-160
  depth regulariser      loss/losses.py:134-148
  geometric consistency  loss/losses.py:84-95
"""
import torch
import torch.nn.functional as F

C1 = 0.01 ** 2
C2 = 0.03 ** 2


def pixel_grid(B, H, W, dtype=torch.float32):
    """[x=col, y=row, 1] rows, flat index j = y*W + x  (view_synthesis.py:17-31)."""
    ys, xs = torch.meshgrid(torch.arange(H, dtype=dtype), torch.arange(W, dtype=dtype), indexing="ij")
    g = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(H * W, dtype=dtype)], 0)
    return g.unsqueeze(0).repeat(B, 1, 1)


def backproject(depth, inv_K):
    B, _, H, W = depth.shape
    grid = pixel_grid(B, H, W, depth.dtype)
    rays = torch.matmul(inv_K[:, :3, :3], grid)                 # :36
    pts = depth.view(B, 1, -1) * rays                           # :38
    return torch.cat([pts, torch.ones(B, 1, H * W, dtype=depth.dtype)], 1)   # :39


def project(points, K, T, H, W, eps=1e-7, geometric=False):
    B = points.shape[0]
    P = torch.matmul(K, T)[:, :3, :]                            # :57
    c = torch.matmul(P, points)                                 # :59
    pix = c[:, :2, :] / (c[:, 2, :].unsqueeze(1) + eps)         # :60
    pix = pix.view(B, 2, H, W).permute(0, 2, 3, 1)              # :61-63
    pix = torch.stack([pix[..., 0] / (W - 1), pix[..., 1] / (H - 1)], -1)   # :66-67 (in-place there)
    pix = (pix - 0.5) * 2                                       # :68
    valid = (pix.abs().max(dim=-1)[0] <= 1).unsqueeze(1).to(points.dtype)   # :70-71
    if geometric:
        wdepth = c[:, 2].clamp(min=1e-3).reshape(B, 1, H, W)    # :74-75
        return pix, wdepth, valid
    return pix, valid


def ssim(x, y, return_raw=False):
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")                  # losses.py:24-25
    y = F.pad(y, (1, 1, 1, 1), mode="reflect")
    mu_x = F.avg_pool2d(x, 3, 1)                                # :27-28
    mu_y = F.avg_pool2d(y, 3, 1)
    sig_x = F.avg_pool2d(x ** 2, 3, 1) - mu_x ** 2              # :30-32
    sig_y = F.avg_pool2d(y ** 2, 3, 1) - mu_y ** 2
    sig_xy = F.avg_pool2d(x * y, 3, 1) - mu_x * mu_y
    n = (2 * mu_x * mu_y + C1) * (2 * sig_xy + C2)              # :34
    d = (mu_x ** 2 + mu_y ** 2 + C1) * (sig_x + sig_y + C2)     # :35
    raw = (1 - n / d) / 2
    if return_raw:                                              # the value BEFORE the clamp (tests use it to find clamp kinks)
        return raw
    return torch.clamp(raw, 0, 1)                               # :37


def photometric(pred, target):
    s = ssim(pred, target).mean(1, True)                        # losses.py:111
    l1 = torch.abs(target - pred).mean(1, True)                 # :112-113
    return 0.85 * s + 0.15 * l1                                 # :115


def warp_photometric(depth, inv_K, K, T, src, tgt, padding_mode="border", use_mask=True, eps=1e-7):
    """One source frame: returns (loss_map[B,1,H,W], syn[B,3,H,W], valid[B,1,H,W], pix[B,H,W,2])."""
    B, _, H, W = depth.shape
    pts = backproject(depth, inv_K)
    pix, valid = project(pts, K, T, H, W, eps)
    syn = F.grid_sample(src, pix, padding_mode=padding_mode, align_corners=False)
    if use_mask:
        pred, target = syn * valid, tgt * valid                 # train_depth.py:714-715
    else:
        pred, target = syn, tgt
    return photometric(pred, target), syn, valid, pix


def photometric_total(depth, inv_K, K, Ts, srcs, tgt, padding_mode="border", use_mask=True,
                      min_reprojection=False):
    """All source frames -> scalar (train_depth.py:726, 629, 657-660)."""
    maps = [warp_photometric(depth, inv_K, K, T, s, tgt, padding_mode, use_mask)[0] for T, s in zip(Ts, srcs)]
    maps = torch.cat(maps, 1)
    if min_reprojection and maps.shape[1] > 1:
        return torch.min(maps, dim=1)[0].mean()
    return maps.mean(1, keepdim=True).mean()


def photometric_objective(depth, inv_K, K, Ts, srcs, tgt, padding_mode="border", use_mask=True, min_reprojection=False,
                          auto_masking=False, noise=None):
    """compute_losses with compute_photometric_loss / compute_automasking_loss (train_depth.py:615-660, 707-750) for S source frames:
    returns (scalar, index of the winning candidate per pixel or None).  `noise` (B,S,H,W) is the tie-breaking term of :646."""
    maps, ident = [], []
    for T, s in zip(Ts, srcs):
        lm, _, valid, _ = warp_photometric(depth, inv_K, K, T, s, tgt, padding_mode, use_mask)
        maps.append(lm)
        ident.append(photometric(s * valid, tgt * valid) if use_mask else photometric(s, tgt))     # :736-747
    photo = torch.cat(maps, 1)                                  # :726
    if not min_reprojection:
        photo = photo.mean(1, keepdim=True)                     # :629
    if auto_masking:
        am = torch.cat(ident, 1)                                # :749
        am = am + noise if min_reprojection else am.mean(1, keepdim=True)     # :645-649
        photo = torch.cat((am, photo), dim=1)                   # :651
    if photo.shape[1] == 1:
        return photo.mean(), None                               # :653-655
    v, i = torch.min(photo, dim=1)                              # :657
    return v.mean(), i


def smoothness(disp, img):
    """compute_smoothness_loss (train_depth.py:767-771) + disparity_smoothness_loss (losses.py:119-132)."""
    m = disp.mean(2, True).mean(3, True)
    n = disp / (m + 1e-7)
    gdx = torch.abs(n[:, :, :, :-1] - n[:, :, :, 1:])
    gdy = torch.abs(n[:, :, :-1, :] - n[:, :, 1:, :])
    gix = torch.mean(torch.abs(img[:, :, :, :-1] - img[:, :, :, 1:]), 1, keepdim=True)
    giy = torch.mean(torch.abs(img[:, :, :-1, :] - img[:, :, 1:, :]), 1, keepdim=True)
    gdx = gdx * torch.exp(-gix)
    gdy = gdy * torch.exp(-giy)
    return gdx.mean() + gdy.mean()


def sparse_gt_l1(pred, sparse_gt, sparse_mask):
    """depth_gt_loss (losses.py:151-160): mean over ALL pixels of |pred*m - gt|."""
    return torch.mean(torch.abs(pred.squeeze() * sparse_mask.squeeze() - sparse_gt.squeeze()))


def depth_regulariser(initial, refined, kind):
    if kind == "l1":
        return torch.mean(torch.abs(initial - refined))
    if kind == "l2":
        return torch.mean((initial - refined) ** 2)
    raise ValueError("please specify a correct norm")


def geometric_consistency(warped_depth, interp_depth, valid):
    diff = ((warped_depth - interp_depth).abs() / (warped_depth + interp_depth)).clamp(0, 1)
    mask = valid.expand_as(diff)
    if mask.sum() > 10000:
        return (diff * mask).sum() / mask.sum()
    return torch.zeros((), dtype=diff.dtype)


def fwd_bwd(depth, inv_K, K, T, src, tgt, padding_mode="border", use_mask=True, dtype=torch.float32,
            want=("depth", "src", "T")):
    """Forward + autograd backward of mean(loss_map).  Inputs are detached/cloned first.

    `src`/`tgt` are given channels-last (B,H,W,3) like the reference's data loader provides them;
    they are viewed NCHW exactly as train_depth.py:451-453 does."""
    depth = depth.detach().to(dtype).clone().requires_grad_("depth" in want)
    T = T.detach().to(dtype).clone().requires_grad_("T" in want)
    src_cl = src.detach().to(dtype).clone().requires_grad_("src" in want)
    tgt_cl = tgt.detach().to(dtype)
    lm, syn, valid, pix = warp_photometric(depth, inv_K.to(dtype), K.to(dtype), T,
                                           src_cl.permute(0, 3, 1, 2), tgt_cl.permute(0, 3, 1, 2),
                                           padding_mode, use_mask)
    loss = lm.mean()
    loss.backward()
    return dict(loss=loss.detach(), loss_map=lm.detach(), syn=syn.detach(), valid=valid.detach(), pix=pix.detach(),
                g_depth=depth.grad, g_src=src_cl.grad, g_T=T.grad)
