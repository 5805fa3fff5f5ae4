"""BASELINE.json configs that are parity cases rather than bench lines:

  C5  4-scale pyramid, 2 source frames per target, 1080x1920 -- the reference modules instantiated once per scale
      (the reference itself only ever runs scale 0, train_depth.py:269);
  C2  TUM-shaped pair with depth holes, larger motion, 3 refinement steps on photometric + point supervision
      (weight 1.0) + depth smoothness (1e-3) + sparse depth supervision (p = 0.012).

Both run the B200 kernels through the public Python surface (ctypes -> C ABI) and compare with the torch-op oracle
(bit-identical to the imported reference on the committed goldens, tests/test_oracle_golden.py) on the same inputs.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_max, same_values
from test_warp_photo_gpu import RTOL, assert_grad_close

pytestmark = pytest.mark.gpu


def _pyramid_level(d, s):
    """Scale s of the pyramid: area-downsampled depth / images, K rows 0-1 scaled by 2^-s."""
    if s == 0:
        return d
    f = 2 ** s
    B, L, H, W, _ = d["colors"].shape
    colors = F.avg_pool2d(d["colors"].permute(0, 1, 4, 2, 3).reshape(B * L, 3, H, W), f)
    colors = colors.reshape(B, L, 3, H // f, W // f).permute(0, 1, 3, 4, 2).contiguous()
    depth = F.avg_pool2d(d["depth"], f)
    K = d["K"].clone()
    K[:, :2, :] = K[:, :2, :] / f
    out = dict(d)
    out.update(colors=colors, depth=depth, K=K, inv_K=torch.pinverse(K))
    return out


@pytest.mark.parametrize("scale", [3, 2, 1, 0])
def test_c5_pyramid_two_sources(scale):
    """Per scale: the loss maps of both source frames are bit-exact, the frame-averaged scalar loss
    (train_depth.py:726, 629, 657) is within 1e-5, and the accumulated depth gradient plus the per-frame source /
    pose gradients meet the gradient bar."""
    import e2e_slam_b200 as e2e
    from e2e_slam_b200.synthetic import make_pairs, se3_exp
    from oracle import torch_oracle as to
    d0 = make_pairs(1, 1080, 1920, "icl", seed=5, rot_deg=1.5, trans=0.04, frames=3)
    d = _pyramid_level(d0, scale)
    T2 = se3_exp(torch.tensor([[0.004, -0.011, 0.006]]), torch.tensor([[-0.03, 0.01, 0.02]]))
    Ts = [d["T"], T2]
    tgt_cl = d["colors"][:, 0]
    srcs_cl = [d["colors"][:, 1], d["colors"][:, 2]]
    H, W = d["depth"].shape[2:]
    assert (H, W) == (1080 >> scale, 1920 >> scale)

    # ---- oracle (CPU): per source frame forward + autograd, fp32 and fp64 ---------------------------------
    ref32 = [to.fwd_bwd(d["depth"], d["inv_K"], d["K"], T, s, tgt_cl, "border", True) for T, s in zip(Ts, srcs_cl)]
    ref64 = [to.fwd_bwd(d["depth"], d["inv_K"], d["K"], T, s, tgt_cl, "border", True, dtype=torch.float64)
             for T, s in zip(Ts, srcs_cl)]
    loss_ref = float(torch.cat([r["loss_map"] for r in ref32], 1).mean(1, keepdim=True).mean())

    # ---- ours ---------------------------------------------------------------------------------------------
    cu = {k: v.cuda() for k, v in d.items()}
    depth = cu["depth"].clone().requires_grad_(True)
    tgt = cu["colors"][:, 0].permute(0, 3, 1, 2)
    srcs = [cu["colors"][:, 1 + i].permute(0, 3, 1, 2).detach().requires_grad_(True) for i in range(2)]
    Tc = [T.cuda().clone().requires_grad_(True) for T in Ts]
    with torch.no_grad():
        for i in range(2):
            lm = e2e.warp_photometric(cu["depth"], cu["inv_K"], cu["K"], Tc[i], srcs[i], tgt, "border", True)
            assert same_values(lm.cpu().numpy(), ref32[i]["loss_map"].numpy()) == 0, f"loss map of source {i} differs"
    total = sum(e2e.warp_photometric_loss(depth, cu["inv_K"], cu["K"], Tc[i], srcs[i], tgt, "border", True) for i in range(2)) / 2
    total.backward()
    assert abs(float(total) - loss_ref) <= RTOL * abs(loss_ref)
    g = lambda rs, k: sum(r[k] for r in rs).numpy() / 2
    assert_grad_close("grad_depth", depth.grad.cpu().numpy(), g(ref32, "g_depth"), g(ref64, "g_depth"))
    for i in range(2):
        assert_grad_close(f"grad_src[{i}]", srcs[i].grad.permute(0, 2, 3, 1).cpu().numpy(), ref32[i]["g_src"].numpy() / 2,
                          ref64[i]["g_src"].numpy() / 2)
        assert_grad_close(f"grad_T[{i}]", Tc[i].grad[:, :3].cpu().numpy(), ref32[i]["g_T"][:, :3].numpy() / 2,
                          ref64[i]["g_T"][:, :3].numpy() / 2)


def _c2_inputs():
    from e2e_slam_b200.synthetic import make_pairs
    d = make_pairs(1, 480, 640, "tum", seed=17, rot_deg=5.0, trans=0.15, holes=0.0)
    g = torch.Generator().manual_seed(3)
    gt_depth = d["depth"].clone()
    holes = torch.rand(1, 1, 480, 640, generator=g) < 0.15                       # 15 % of the GT depth missing
    gt_depth[holes] = 0.0
    disp0 = 1.0 / (d["depth"] * (1.0 + 0.05 * torch.randn(1, 1, 480, 640, generator=g)))   # noisy prediction (disparity)
    mask = ((torch.rand(1, 480, 640, 1, generator=g) < 0.012) & (gt_depth.permute(0, 2, 3, 1) != 0)).float()   # training_utils.py:176-189
    sparse_gt = gt_depth.permute(0, 2, 3, 1) * mask
    # global map: the GT surface seen from the previous frame, jittered; 30 k points
    idx = torch.randperm(480 * 640, generator=g)[:30000]
    ys, xs = (idx // 640).float(), (idx % 640).float()
    z = d["depth"][0, 0].reshape(-1)[idx]
    K = d["K"][0]
    pts = torch.stack([(xs - K[0, 2]) / K[0, 0] * z, (ys - K[1, 2]) / K[1, 1] * z, z], 1)
    pts = pts + 0.002 * torch.randn(pts.shape, generator=g)
    return d, disp0, sparse_gt, mask, pts.contiguous()


def _c2_step_oracle(disp, d, sparse_gt, mask, gmap, sub):
    """One refinement step's loss in plain torch on the CPU: the reference's composition (train_depth.py:615-705,
    online_adaption.py:638-645) with the oracle's restatements."""
    from oracle import fusion_oracle as fo
    from oracle import torch_oracle as to
    depth = 1.0 / disp
    src, tgt = d["colors"][:, 0].permute(0, 3, 1, 2), d["colors"][:, 1].permute(0, 3, 1, 2)
    photo = to.warp_photometric(depth, d["inv_K"], d["K"], d["T"], src, tgt, "border", True)[0].mean()
    smooth = to.smoothness(disp, tgt)
    gt_l1 = to.sparse_gt_l1(depth, sparse_gt, mask)
    # live cloud (a fixed pixel subset keeps the CPU kNN in seconds) moved by T, nearest map point, detached indices
    cam = to.backproject(depth, d["inv_K"])[0, :3].t()[sub]
    q = cam @ d["T"][0, :3, :3].t() + d["T"][0, :3, 3]
    _, idx = fo.knn1(q.detach().numpy(), gmap.numpy(), chunk=1024)
    knn = ((q - gmap[torch.from_numpy(idx)]) ** 2).sum(1).mean()
    return photo + 1.0 * knn + 1e-3 * smooth + gt_l1, dict(photo=float(photo), knn=float(knn), smooth=float(smooth), gt=float(gt_l1))


def _c2_step_ours(disp, cu, sparse_gt, mask, gmap, sub):
    import e2e_slam_b200 as e2e
    from e2e_slam_b200 import losses, view_synthesis
    depth = 1.0 / disp
    src, tgt = cu["colors"][:, 0].permute(0, 3, 1, 2), cu["colors"][:, 1].permute(0, 3, 1, 2)
    photo = e2e.warp_photometric_loss(depth, cu["inv_K"], cu["K"], cu["T"], src, tgt, "border", True)
    smooth = losses.smoothness_loss(disp, tgt)
    gt_l1 = losses.depth_gt_loss(depth, sparse_gt, mask)
    cam = view_synthesis.BackprojectDepth(1, 480, 640)(depth, cu["inv_K"])[0, :3].t()[sub]
    knn = losses.point_supervision_loss(cam, cu["T"][0], gmap)
    return photo + 1.0 * knn + 1e-3 * smooth + gt_l1, dict(photo=float(photo), knn=float(knn), smooth=float(smooth), gt=float(gt_l1))


def test_c2_three_refinement_steps():
    """Three gradient steps on the predicted disparity under the C2 loss mix; every loss term of every step and the
    refined disparity must follow the CPU oracle's trajectory."""
    d, disp0, sparse_gt, mask, gmap = _c2_inputs()
    sub = torch.randperm(480 * 640, generator=torch.Generator().manual_seed(9))[:8000]
    cu = {k: v.cuda() for k, v in d.items()}
    lr = 5e-3
    disp_ref, disp_gpu = disp0.clone(), disp0.clone().cuda()
    for step in range(3):
        a = disp_ref.clone().requires_grad_(True)
        loss_ref, terms_ref = _c2_step_oracle(a, d, sparse_gt, mask, gmap, sub)
        loss_ref.backward()
        b = disp_gpu.clone().requires_grad_(True)
        loss, terms = _c2_step_ours(b, cu, sparse_gt.cuda(), mask.cuda(), gmap.cuda(), sub.cuda())
        loss.backward()
        for k in terms_ref:
            assert abs(terms[k] - terms_ref[k]) <= 2e-5 * max(abs(terms_ref[k]), 1e-6), f"step {step}: {k} {terms[k]} vs {terms_ref[k]}"
        gr, go = a.grad.numpy(), b.grad.cpu().numpy()
        l2 = float(np.linalg.norm(go - gr) / np.linalg.norm(gr))
        assert l2 <= 5e-5, f"step {step}: disparity gradient relative L2 error {l2:.2e}"
        disp_ref = (disp_ref - lr * a.grad).detach()
        disp_gpu = (disp_gpu - lr * b.grad).detach()
        assert rel_max(disp_gpu.cpu().numpy(), disp_ref.numpy()) <= 1e-5
    assert float((disp_ref - disp0).abs().max()) > 1e-6        # the steps really moved the prediction
