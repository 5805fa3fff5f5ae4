"""GPU parity tests of the entry points added in round 2: find_active_map_points rows, transform_pointcloud, color_points_loss,
ChamferDistance (forward / reverse / bidirectional), edge cases of the K = 1 nearest-neighbour op.  Oracles: oracle/fusion_oracle.py
(numpy; gradslam / chamferdist semantics as frozen there -- parity unpinned) and plain torch CPU ops."""
import numpy as np
import pytest
import torch

from conftest import rel_max, same_values

pytestmark = pytest.mark.gpu


def _room_map(L, H, W, seed=0):
    from oracle import fusion_oracle as fo
    depth, rgb, K, poses = fo.synthetic_room_sequence(L, H, W, seed)
    o = fo.PointFusionOracle(0.05, 20, 0.6)
    for s in range(L - 1):
        o.step(depth[s].astype(np.float32), rgb[s].astype(np.float32), K, poses[s])
    return o, depth.astype(np.float32), rgb.astype(np.float32), K, poses


@pytest.mark.parametrize("H,W,B", [(60, 80, 1), (120, 160, 2)])
def test_find_active_map_points_rows_bit_exact(H, W, B):
    """gradslam.slam.fusionutils.find_active_map_points (online_adaption.py:35): (N_active, 4) int64 rows (b, n, h, w) ordered by (b, n),
    bit-exact against the oracle's step 1, through the patched import path the reference script would take."""
    import e2e_slam_b200.patch as patch
    from e2e_slam_b200.slam import Pointclouds, RGBDImages
    from oracle import fusion_oracle as fo
    patch.install()
    from gradslam.slam.fusionutils import find_active_map_points
    L = 4
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    pts, nrm, col, cc, rgbs, depths, Ks, ps, expect = [], [], [], [], [], [], [], [], []
    for b in range(B):
        o, depth, rgb, K, poses = _room_map(L, H, W, seed=b)
        pts.append(t(o.points)); nrm.append(t(o.normals)); col.append(t(o.colors)); cc.append(t(o.ccount))
        rgbs.append(t(rgb[L - 1])[None]); depths.append(t(depth[L - 1])[None, ..., None]); Ks.append(t(K).view(1, 4, 4)); ps.append(t(poses[L - 1])[None])
        expect.append(fo.find_active_map_points(o.points, K, poses[L - 1], H, W, b))
    pc = Pointclouds(points=pts, normals=nrm, colors=col, features=cc, device="cuda")
    live = RGBDImages(torch.stack(rgbs), torch.stack(depths), torch.stack(Ks), torch.stack(ps))
    rows = find_active_map_points(pc, live)
    expect = np.concatenate(expect, 0)
    assert rows.dtype == torch.int64 and rows.shape == expect.shape and len(expect) > 100
    assert np.array_equal(rows.cpu().numpy(), expect)
    # empty map -> no rows; wrong sequence length -> ValueError like gradslam
    assert find_active_map_points(Pointclouds(points=[p[:0] for p in pts], device="cuda"), live).shape == (0, 4)
    with pytest.raises(ValueError):
        find_active_map_points(pc, RGBDImages(torch.cat([live.rgb_image] * 2, 1), torch.cat([live.depth_image] * 2, 1), live.intrinsics,
                                              torch.cat([live.poses] * 2, 1)))


def test_transform_pointcloud_bit_exact_and_grads():
    from e2e_slam_b200.slam import transform_pointcloud
    from e2e_slam_b200.synthetic import se3_exp
    from oracle import fusion_oracle as fo
    g = torch.Generator().manual_seed(1)
    p = torch.randn(5001, 3, generator=g) * 2
    T = se3_exp(torch.tensor([[0.03, -0.02, 0.05]]), torch.tensor([[0.1, -0.2, 0.05]]))[0]
    pd, Td = p.cuda().requires_grad_(True), T.cuda().requires_grad_(True)
    out = transform_pointcloud(pd, Td)
    assert same_values(out.detach().cpu().numpy(), fo.transform_pointcloud(p.numpy(), T.numpy())) == 0
    w = torch.randn(5001, 3, generator=g)
    (out * w.cuda()).sum().backward()
    p64, T64 = p.double().requires_grad_(True), T.double().requires_grad_(True)
    ((p64 @ T64[:3, :3].t() + T64[:3, 3]) * w.double()).sum().backward()
    assert rel_max(pd.grad.cpu().numpy(), p64.grad.numpy()) <= 1e-6
    assert rel_max(Td.grad.cpu().numpy(), T64.grad.numpy()) <= 1e-5
    assert transform_pointcloud(pd[:0], Td).shape == (0, 3)
    with pytest.raises(ValueError):
        transform_pointcloud(pd[None], Td)


def test_color_points_loss_matches_reference_expression():
    """loss/losses.py:65-82: mean |noisy_col[0] - gt_col[0, idx[0]]|, and its gradients to both colour clouds."""
    from e2e_slam_b200.losses import color_points_loss, knn_points_loss
    g = torch.Generator().manual_seed(2)
    P1, P2 = 7001, 3003
    gt_p, no_p = torch.rand(1, P2, 3, generator=g), torch.rand(1, P1, 3, generator=g)
    gt_c, no_c = torch.rand(1, P2, 3, generator=g), torch.rand(1, P1, 3, generator=g)
    no_c[0, :50] = gt_c[0, :50]                       # exact ties exercise sign(0) = 0 when the match happens to be the same row
    _, idx = knn_points_loss(gt_p.cuda(), no_p.cuda())
    a, b = gt_c.cuda().requires_grad_(True), no_c.cuda().requires_grad_(True)
    loss = color_points_loss(a, b, idx)
    (loss * 1.7).backward()
    a64, b64 = gt_c.double().requires_grad_(True), no_c.double().requires_grad_(True)
    ref = torch.mean(torch.abs(b64[0] - a64[0, idx[0].cpu()]))
    (ref * 1.7).backward()
    assert abs(float(loss) - float(ref)) <= 1e-6 * float(ref)
    assert rel_max(b.grad.cpu().numpy(), b64.grad.numpy()) <= 1e-6
    assert rel_max(a.grad.cpu().numpy(), a64.grad.numpy()) <= 1e-5
    with pytest.raises(ValueError):
        color_points_loss(a[..., :2], b, idx)


def test_chamfer_distance_directions():
    """chamferdist.ChamferDistance as train_depth.py:689-695 would call it: forward, reverse and bidirectional sums of K = 1 nearest-
    neighbour distances; distances bit-exact against the brute-force oracle, reductions to 1e-6, gradients against float64."""
    from e2e_slam_b200.losses import ChamferDistance
    from oracle import fusion_oracle as fo
    g = torch.Generator().manual_seed(3)
    src, tgt = torch.rand(1, 4099, 3, generator=g), torch.rand(1, 2500, 3, generator=g) + 0.05
    d_st, i_st = fo.knn1(src[0].numpy(), tgt[0].numpy())
    d_ts, i_ts = fo.knn1(tgt[0].numpy(), src[0].numpy())
    cd = ChamferDistance()
    for kw, ref in ((dict(), d_st.astype(np.float64).mean()), (dict(reverse=True), d_ts.astype(np.float64).mean()),
                    (dict(bidirectional=True), d_st.astype(np.float64).mean() + d_ts.astype(np.float64).mean()),
                    (dict(bidirectional=True, reduction="sum"), d_st.astype(np.float64).sum() + d_ts.astype(np.float64).sum())):
        s, t = src.cuda().requires_grad_(True), tgt.cuda().requires_grad_(True)
        v = cd(s, t, **kw)
        assert abs(float(v) - ref) <= 2e-6 * ref, kw
        v.backward()
        s64, t64 = src.double().requires_grad_(True), tgt.double().requires_grad_(True)
        f = ((s64[0] - t64[0][torch.from_numpy(i_st)]) ** 2).sum(1)
        r = ((t64[0] - s64[0][torch.from_numpy(i_ts)]) ** 2).sum(1)
        red = (lambda x: x.sum()) if kw.get("reduction") == "sum" else (lambda x: x.mean())
        tot = red(r) if kw.get("reverse") else (red(f) + red(r) if kw.get("bidirectional") else red(f))
        tot.backward()
        assert rel_max(s.grad.cpu().numpy(), s64.grad.numpy()) <= 1e-5, kw
        assert rel_max(t.grad.cpu().numpy(), t64.grad.numpy()) <= 1e-5, kw
    with pytest.raises(ValueError):
        cd(src.cuda(), tgt.cuda(), reduction="max")


def test_knn_edge_cases():
    """An all-invalid live depth map gives an empty query cloud (online_adaption.py:638-645): empty outputs, not an error; an empty
    reference cloud is an error; the fused transform is not differentiated and says so."""
    from e2e_slam_b200.losses import knn_points, point_supervision_loss
    ref = torch.rand(1, 100, 3).cuda()
    out = knn_points(torch.zeros(1, 0, 3).cuda(), ref)
    assert out.dists.shape == (1, 0, 1) and out.idx.shape == (1, 0, 1) and out.idx.dtype == torch.int64
    with pytest.raises(ValueError):
        knn_points(ref, torch.zeros(1, 0, 3).cuda())
    T = torch.eye(4).cuda().requires_grad_(True)
    with pytest.raises(NotImplementedError):
        point_supervision_loss(ref[0], T, ref[0])


@pytest.mark.parametrize("mr,am", [(False, False), (True, False), (False, True), (True, True)])
def test_multi_source_objective_vs_reference_goldens(composite_golden, mr, am):
    """S = 2 source frames per target through warp_photometric_multi: plain frame mean, min-reprojection, auto-masking and both
    (train_depth.py:615-660, 707-750), against goldens produced with the reference's own modules (tools/make_golden_composite.py).
    The selected candidate per pixel must be identical (the loss maps are bit-exact), the loss and the gradients within 1e-5 of the
    reference's fp32 result or as close to it as the reference is to its float64 evaluation."""
    import e2e_slam_b200 as e2e
    from test_warp_photo_gpu import assert_grad_close
    g = composite_golden
    tag = f"mr{int(mr)}_am{int(am)}"
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    depth = t(g["depth"]).requires_grad_(True)
    colors = t(g["colors"]).requires_grad_(True)
    K, inv_K = t(g["K"]), t(g["inv_K"])            # the reference's own pinverse (train_depth.py:460-461), computed on the CPU
    S = g["T"].shape[1]
    tgt = colors[:, 0].permute(0, 3, 1, 2)
    srcs = [colors[:, 1 + s].permute(0, 3, 1, 2) for s in range(S)]
    Ts = [t(g["T"][:, s]) for s in range(S)]
    loss = e2e.warp_photometric_multi(depth, inv_K, K, Ts, srcs, tgt, str(g["padding_mode"]), bool(g["use_mask"]),
                                      min_reprojection=mr, auto_masking=am, noise=t(g["noise"]))
    loss.backward()
    ref = float(g[f"loss_{tag}"])
    assert abs(float(loss) - ref) <= 1e-5 * abs(ref), (tag, float(loss), ref)
    assert_grad_close("g_depth " + tag, depth.grad.cpu().numpy(), g[f"g_depth_{tag}"], g[f"g_depth_{tag}_f64"])
    # sources only (the reference differentiates the target through the identity / L1 terms; this path does not: SURVEY appendix A)
    assert_grad_close("g_sources " + tag, colors.grad[:, 1:].cpu().numpy(), g[f"g_colors_{tag}"][:, 1:], g[f"g_colors_{tag}_f64"][:, 1:])


def test_min_composite_matches_torch_min_with_ties_and_nan():
    """e2e_min_composite_fwd/bwd against torch.min(dim=1) + mean on CPU: first minimum wins, NaN propagates, gradient to the winner."""
    from e2e_slam_b200.losses import photometric_objective
    g = torch.Generator().manual_seed(0)
    B, H, W, C = 2, 37, 53, 5
    maps = [torch.rand(B, 1, H, W, generator=g) for _ in range(C)]
    maps[3][:, :, :5] = maps[1][:, :, :5]                        # exact ties: the earlier candidate must win
    maps[1][:, :, :5] = maps[1][:, :, :5].clamp(max=0.001)
    maps[3][:, :, :5] = maps[1][:, :, :5]
    dev = [m.cuda().requires_grad_(True) for m in maps]
    loss, index = photometric_objective(dev, None, min_reprojection=True, return_index=True)
    (loss * 3.0).backward()
    cpu = [m.clone().requires_grad_(True) for m in maps]
    v, i = torch.min(torch.cat(cpu, 1), dim=1)
    (v.mean() * 3.0).backward()
    assert np.array_equal(index[:, 0].cpu().numpy(), i.numpy())
    assert abs(float(loss) - float(v.mean())) <= 1e-6 * float(v.mean())
    for a, b in zip(dev, cpu):
        ga, gb = a.grad.cpu().numpy(), b.grad.numpy()
        assert np.array_equal(ga != 0, gb != 0)                    # the gradient goes to the winner and only to the winner
        assert rel_max(ga, gb) <= 1e-6                             # (3 * (1/n) vs 3/n: one rounding apart)
    maps[2][0, 0, 7, 7] = float("nan")
    loss2, index2 = photometric_objective([m.cuda() for m in maps], None, min_reprojection=True, return_index=True)
    assert bool(torch.isnan(loss2)) and int(index2[0, 0, 7, 7]) == 2


@pytest.mark.parametrize("B,L,H,W", [(3, 8, 60, 80), (4, 6, 120, 160)])
def test_batched_fusion_equals_single_sequences(B, L, H, W):
    """PointFusion.__call__ on a batch of B different sequences (gradslam's batch dimension, train_depth.py:263-267) runs ONE cooperative
    launch over all of them (e2e_fusion_sequence_batch); every sequence's map must be, bit for bit, the one the single-sequence path
    builds for it -- and that path is bit-exact against the oracle (test_fusion_gpu.py)."""
    from e2e_slam_b200.slam import PointFusion, RGBDImages
    from oracle import fusion_oracle as fo
    seqs = []
    for b in range(B):
        depth, rgb, K, poses = fo.synthetic_room_sequence(L, H, W, seed=10 + b)
        depth = (depth * (np.random.default_rng(b).random(depth.shape) >= 0.05 * b)).astype(np.float32)
        seqs.append((depth, rgb.astype(np.float32), K, poses))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    batch = RGBDImages(torch.stack([t(s[1]) for s in seqs]), torch.stack([t(s[0]) for s in seqs])[..., None],
                       torch.stack([t(s[2]) for s in seqs])[:, None], torch.stack([t(s[3]) for s in seqs]))
    slam = PointFusion(odom="gt", dist_th=0.05, angle_th=20, sigma=0.6, device="cuda")
    with torch.no_grad():
        pcs, poses_out = slam(batch)
        assert len(pcs) == B and torch.equal(poses_out, batch.poses)
        for b in range(B):
            single, _ = slam(batch[b])
            for name in ("points_list", "normals_list", "colors_list", "features_list"):
                a, c = getattr(pcs, name)[b].cpu().numpy(), getattr(single, name)[0].cpu().numpy()
                assert a.shape == c.shape and a.shape[0] > H * W, (b, name, a.shape, c.shape)
                assert same_values(a, c) == 0, (b, name)


def test_median_radix_select_matches_torch():
    """ops.median == torch.median (lower median) bit for bit: random values with ties, negatives, zeros, +-inf; NaN propagates; and the
    median-scaling ratio of online_adaption.py:295 computed from the DISPARITIES equals the reference expression on materialised depths."""
    from e2e_slam_b200 import ops
    g = torch.Generator().manual_seed(0)
    for n in (1, 2, 7, 1000, 307200, 614401):
        x = torch.randn(n, generator=g)
        x[::5] = x[::5].round()                                   # many exact ties
        if n > 10:
            x[3], x[4], x[5], x[6] = 0.0, -0.0, float("inf"), float("-inf")
        assert float(ops.median(x.cuda())) == float(torch.median(x)), n
        for k in (0, n // 3, n - 1):
            assert float(ops.select_kth(x.cuda(), k)) == float(torch.sort(x)[0][k]), (n, k)
    x[17] = float("nan")
    assert bool(torch.isnan(ops.median(x.cuda())))
    gt = torch.rand(1, 2, 480, 640, 1, generator=g) * 4
    gt[gt < 0.6] = 0.0                                            # TUM-style holes
    disp = torch.rand(2, 1, 480, 640, generator=g) * 2 + 0.05
    depth_tensor = torch.cat([(1 / disp[i:i + 1]).unsqueeze(1) for i in range(2)], 1).permute(0, 1, 3, 4, 2)     # online_adaption.py:282-293
    ref = torch.median(gt) / torch.median(depth_tensor)
    ours = ops.median_ratio_from_disparity(gt.cuda(), disp.cuda())
    assert float(ours) == float(ref)


@pytest.mark.parametrize("H,W,scaled", [(48, 64, True), (50, 70, False), (480, 640, True)])
def test_disparity_fed_sweep_equals_conversion_then_sweep(H, W, scaled):
    """warp_photometric_loss_from_disparity(disp, ratio) folds depth = (1/disp)*ratio into the sweep's depth load:
    the loss must be bit-identical to ops.disp_to_depth + warp_photometric_loss (both 16-byte-aligned TMA and generic-stride instances),
    the disparity gradient equal to the chained one."""
    import e2e_slam_b200 as e2e
    from e2e_slam_b200 import ops
    from e2e_slam_b200.synthetic import make_pairs
    d = make_pairs(2, H, W, "tum", seed=H, rot_deg=3.0, trans=0.1)
    cu = {k: v.cuda() for k, v in d.items()}
    src, tgt = cu["colors"][:, 0].permute(0, 3, 1, 2), cu["colors"][:, 1].permute(0, 3, 1, 2)
    disp0 = (1.0 / cu["depth"]) * 1.3
    ratio = torch.tensor(1.3, device="cuda") if scaled else None
    a = disp0.clone().requires_grad_(True)
    la = e2e.warp_photometric_loss(ops.disp_to_depth(a, ratio), cu["inv_K"], cu["K"], cu["T"], src, tgt, "border", True)
    la.backward()
    b = disp0.clone().requires_grad_(True)
    lb = ops.warp_photometric_loss_from_disparity(b, cu["inv_K"], cu["K"], cu["T"], src, tgt, ratio, "border", True)
    lb.backward()
    assert float(la) == float(lb)
    assert rel_max(b.grad.cpu().numpy(), a.grad.cpu().numpy()) <= 2e-6


class _Args:
    """The slice of the reference's yaml config that the patched methods read (configs/config.yaml)."""

    def __init__(self, min_reprojection=False, auto_masking=False):
        ns = lambda **k: type("NS", (), k)()
        self.DATA = ns(frames=[0, -1, 1])
        self.MODEL = ns(padding_mode="border")
        self.LOSS = ns(geometric=False, photometric_mask=True, min_reprojection=min_reprojection, auto_masking=auto_masking, smoothness=True,
                       smoothness_weight=1e-3, depth_regularizer=False, knn_points=False, chamfer_distance=False, supervise_depth=True,
                       gt_depth_weight=1.0)


def _driver(args, disp, opt):
    """A stand-in for train_depth.Depth_Estimation / online_adaption.SLAM carrying the methods patch.fuse() swaps or calls; the
    un-swapped ones are written as the reference has them (train_depth.py:729-750, 763-796) on top of the patched `loss.losses`."""
    from e2e_slam_b200 import losses, patch

    class Driver:
        def compute_automasking_loss(self, inputs, outputs):                       # train_depth.py:729-750
            out = []
            for frame in self.args.DATA.frames[1:]:
                out.append(losses.photometric_loss(ssim=self.ssim, prediction=inputs[("source_frame", frame)] * outputs["valid_mask", frame],
                                                   target=inputs["target_frame"] * outputs["valid_mask", frame]))
            return torch.cat(out, 1)

        def compute_smoothness_loss(self, inputs):                                  # :763-773
            d = inputs["disp"]
            norm = d / (d.mean(2, True).mean(3, True) + 1e-7)
            return losses.disparity_smoothness_loss(norm, inputs["target_frame"])

        def compute_gt_depth_loss(self, inputs):                                    # :788-796
            return losses.depth_gt_loss(inputs["target_depth"], inputs["sparse_gt"], inputs["sparse_mask"])

    patch.fuse(Driver, defer_items=True, graph=True)
    d = Driver()
    d.args, d.ssim, d.optimizer = args, losses.SSIM(), opt
    return d


@pytest.mark.parametrize("mr,am", [(False, False), (True, True)])
def test_patched_compute_losses_one_read_per_step(mr, am):
    """patch.fuse(cls): novel_view_synthesis + compute_photometric_loss + compute_losses under the reference's method names; the step's
    loss terms must equal the reference composition (torch oracle on the CPU) and the optimizer must have stepped."""
    from e2e_slam_b200.synthetic import make_pairs
    from oracle import torch_oracle as to
    H, W = 48, 64
    d = make_pairs(1, H, W, "tum", seed=5, rot_deg=3.0, trans=0.1, frames=3)
    d2 = make_pairs(1, H, W, "tum", seed=6, rot_deg=3.0, trans=0.1)
    g = torch.Generator().manual_seed(1)
    disp0 = 1.0 / d["depth"] * (1 + 0.05 * torch.randn(1, 1, H, W, generator=g))
    mask = (torch.rand(1, H, W, 1, generator=g) < 0.3).float()
    sparse_gt = d["depth"].permute(0, 2, 3, 1) * mask
    Ts = [d["T"], torch.linalg.inv(d2["T"])]
    cu = lambda t: t.cuda()
    disp = cu(disp0).requires_grad_(True)
    opt = torch.optim.SGD([disp], lr=1e-2)
    drv = _driver(_Args(mr, am), disp, opt)
    colors = cu(d["colors"])
    inputs = {"target_depth": 1.0 / disp, "disp": disp, "Inverse_K": cu(d["inv_K"]), "K": cu(d["K"]), "target_frame": colors[:, 1].permute(0, 3, 1, 2),
              ("source_frame", -1): colors[:, 0].permute(0, 3, 1, 2), ("source_frame", 1): colors[:, 2].permute(0, 3, 1, 2),
              ("T", -1): cu(Ts[0]), ("T", 1): cu(Ts[1]), "sparse_gt": cu(sparse_gt), "sparse_mask": cu(mask)}
    torch.manual_seed(0)
    noise_ref = None
    outputs = drv.novel_view_synthesis(inputs)
    if mr and am:                                   # the tie-breaking noise is drawn inside (train_depth.py:646): too small (1e-5) to matter at 1e-4
        pass
    total = drv.compute_losses(inputs, outputs)
    # ---- reference composition on the CPU
    a = disp0.clone().requires_grad_(True)
    depth = 1.0 / a
    tgt = d["colors"][:, 1].permute(0, 3, 1, 2)
    srcs = [d["colors"][:, 0].permute(0, 3, 1, 2), d["colors"][:, 2].permute(0, 3, 1, 2)]
    photo, _ = to.photometric_objective(depth, d["inv_K"], d["K"], Ts, srcs, tgt, "border", True, mr, am, torch.zeros(1, 2, H, W))
    smooth = to.smoothness(a, tgt)
    gt_l1 = to.sparse_gt_l1(depth, sparse_gt, mask)
    ref = photo + 1e-3 * smooth + gt_l1
    ref.backward()
    tol = 2e-4 if (mr and am) else 1e-5             # with noise: a handful of near-tied pixels may pick the other candidate
    assert abs(total - float(ref)) <= tol * abs(float(ref))
    assert set(drv.last_losses) == {"photometric_loss", "smoothn_loss", "gt_depth_loss"}
    assert abs(drv.last_losses["photometric_loss"] - float(photo)) <= tol * float(photo)
    assert abs(drv.last_losses["smoothn_loss"] - float(smooth)) <= 1e-5 * float(smooth)
    assert abs(drv.last_losses["gt_depth_loss"] - float(gt_l1)) <= 1e-5 * float(gt_l1)
    # optimizer.step() ran, on the right gradient (compared on the updated parameter: the update itself is ~1e-5 of a ~0.5 value)
    assert rel_max(disp.detach().cpu().numpy(), (disp0 - 1e-2 * a.grad).numpy()) <= 1e-6
    assert float((disp.detach().cpu() - disp0).abs().max()) > 1e-6


def test_graphed_refinement_step_matches_eager():
    """patch.GraphedStep: a whole refinement step (photometric + smoothness + sparse depth, backward, Adam) captured as one CUDA graph
    must produce the eager step's losses and parameters, step after step."""
    import e2e_slam_b200 as e2e
    from e2e_slam_b200 import losses, patch
    from e2e_slam_b200.synthetic import make_pairs
    H, W = 96, 128
    d = {k: v.cuda() for k, v in make_pairs(1, H, W, "tum", seed=9, rot_deg=4.0, trans=0.1).items()}
    src, tgt = d["colors"][:, 0].permute(0, 3, 1, 2), d["colors"][:, 1].permute(0, 3, 1, 2)
    g = torch.Generator(device="cuda").manual_seed(2)
    disp0 = 1.0 / d["depth"] * (1 + 0.05 * torch.randn(1, 1, H, W, generator=g, device="cuda"))
    mask = (torch.rand(1, H, W, 1, generator=g, device="cuda") < 0.3).float()
    sparse_gt = d["depth"].permute(0, 2, 3, 1) * mask

    def make(disp):
        opt = torch.optim.Adam([disp], lr=1e-3, capturable=True)

        def step():
            opt.zero_grad(set_to_none=False)
            photo = e2e.ops.warp_photometric_loss_from_disparity(disp, d["inv_K"], d["K"], d["T"], src, tgt, None, "border", True)
            smooth = losses.smoothness_loss(disp, tgt)
            gt_l1 = losses.depth_gt_loss(1.0 / disp, sparse_gt, mask)
            loss = photo + 1e-3 * smooth + gt_l1
            loss.backward()
            opt.step()
            return {"loss": loss.detach(), "photo": photo.detach()}
        return step

    da, db = disp0.clone().requires_grad_(True), disp0.clone().requires_grad_(True)
    eager, graphed = make(da), patch.GraphedStep(make(db), warmup=2)
    for i in range(6):
        ta, tb = eager(), graphed()
        assert abs(float(ta["loss"]) - float(tb["loss"])) <= 1e-6 * abs(float(ta["loss"])), i
        assert rel_max(db.detach().cpu().numpy(), da.detach().cpu().numpy()) <= 1e-6, i
    assert graphed.graph is not None                 # steps 3.. were graph replays
    assert np.isfinite(float(ta["loss"]))


@pytest.mark.parametrize("B,H,W", [(3, 2, 2), (1, 3, 3), (1, 17, 5), (2, 48, 64), (1, 37, 52), (5, 96, 128)])
def test_sweep_writes_stay_inside_its_buffers(B, H, W):
    """compute-sanitizer is closed on this GPU pool, so out-of-bounds WRITES of the streaming kernel (TMA and generic instances, tiny
    and ragged shapes, row segments) are caught with canaries: every output / workspace buffer of WarpPhotoPlan is carved out of one
    arena with 16 KB guard bands on both sides, and the bands must be untouched afterwards.  (A 2-row image once wrote the next pair's
    row 0 through exactly this kind of slip.)"""
    from e2e_slam_b200 import ops
    from e2e_slam_b200.synthetic import make_pairs
    d = {k: v.cuda() for k, v in make_pairs(B, H, W, "tum", seed=W, rot_deg=6.0, trans=0.3).items()}
    src, tgt = d["colors"][:, 0].permute(0, 3, 1, 2), d["colors"][:, 1].permute(0, 3, 1, 2)
    plan = ops.WarpPhotoPlan(B, H, W, "cuda")
    G = 4096                                                     # guard band, in floats
    sizes = {"loss": 1, "grad_depth": B * H * W, "grad_src": B * 3 * H * W, "grad_P": B * 12, "vg_ws": (plan.vg_ws_bytes + 3) // 4}
    total = sum(G + ((n + 63) // 64) * 64 for n in sizes.values()) + G
    arena = torch.full((total,), -7.25, device="cuda")
    off, views = G, {}
    for k, n in sizes.items():
        views[k] = arena[off:off + n]
        off += ((n + 63) // 64) * 64 + G
    plan.loss, plan.grad_depth = views["loss"], views["grad_depth"].view(B, 1, H, W)
    plan.grad_src, plan.grad_P = views["grad_src"].view(B, 3, H, W), views["grad_P"].view(B, 3, 4)
    plan._gs_strides = ops.strides4(plan.grad_src)
    plan.vg_ws = views["vg_ws"].view(torch.uint8)[:plan.vg_ws_bytes]
    ref = ops.WarpPhotoPlan(B, H, W, "cuda")
    l0, gd0, gs0, gp0 = [t.clone() for t in ref.value_and_grad(d["depth"], d["inv_K"], d["K"], d["T"], src, tgt)]
    l1, gd1, gs1, gp1 = plan.value_and_grad(d["depth"], d["inv_K"], d["K"], d["T"], src, tgt)
    torch.cuda.synchronize()
    assert float(l0) == float(l1) and torch.equal(gd0, gd1) and torch.equal(gp0, gp1)
    mask = torch.ones(total, dtype=torch.bool, device="cuda")
    off = G
    for k, n in sizes.items():
        mask[off:off + n] = False
        off += ((n + 63) // 64) * 64 + G
    assert bool((arena[mask] == -7.25).all()), "a guard band was written"


@pytest.mark.parametrize("B,S,H,W,pad", [(2, 2, 48, 64, "border"), (1, 3, 37, 52, "zeros"), (2, 2, 270, 480, "border")])
def test_multi_source_single_launch_equals_per_frame_sweeps(B, S, H, W, pad):
    """e2e_warp_photo_vg_multi (S source frames per target in one launch): the loss must equal the mean of the S per-frame
    losses (each bit-exact against the oracles elsewhere) and every gradient the sum / slice of the per-frame gradients."""
    import e2e_slam_b200 as e2e
    from e2e_slam_b200 import ops
    from e2e_slam_b200.synthetic import make_pairs
    ds = [make_pairs(B, H, W, "icl", seed=40 + s, rot_deg=3.0, trans=0.1) for s in range(S)]
    cu = {k: v.cuda() for k, v in ds[0].items()}
    tgt = cu["colors"][:, 1].permute(0, 3, 1, 2)
    srcs_cl = torch.stack([d["colors"][:, 0] for d in ds], 1).cuda()                       # (B,S,H,W,3) channels-last, like the loader
    Ts = torch.stack([d["T"] for d in ds], 1).cuda()
    depth_a = cu["depth"].clone().requires_grad_(True)
    src_a, T_a = srcs_cl.clone().requires_grad_(True), Ts.clone().requires_grad_(True)
    la = ops.warp_photometric_loss_multi(depth_a, cu["inv_K"], cu["K"], T_a, src_a.permute(0, 1, 4, 2, 3), tgt, pad, True)
    la.backward()
    depth_b = cu["depth"].clone().requires_grad_(True)
    src_b, T_b = srcs_cl.clone().requires_grad_(True), Ts.clone().requires_grad_(True)
    lb = sum(e2e.warp_photometric_loss(depth_b, cu["inv_K"], cu["K"], T_b[:, s], src_b[:, s].permute(0, 3, 1, 2), tgt, pad, True)
             for s in range(S)) / S
    lb.backward()
    assert abs(float(la) - float(lb)) <= 1e-6 * float(lb), (float(la), float(lb))
    # same kernel arithmetic; the paths differ in the constant factor (1/(B S H W) vs 1/(B H W) then /S) and the atomics' order
    assert rel_max(depth_a.grad.cpu().numpy(), depth_b.grad.cpu().numpy()) <= 1e-5
    assert rel_max(src_a.grad.cpu().numpy(), src_b.grad.cpu().numpy()) <= 1e-5
    assert rel_max(T_a.grad.cpu().numpy(), T_b.grad.cpu().numpy()) <= 5e-5


@pytest.mark.parametrize("B,H,W,disp", [(2, 48, 64, False), (1, 9, 12, False), (3, 120, 160, True), (2, 480, 640, False), (1, 271, 480, False)])
def test_role_split_kernel_equals_classic_kernel(monkeypatch, B, H, W, disp):
    """E2E_ROLES=1 runs the lean value + gradient sweep on warp_photo_roles_kernel (three warp groups, progress mbarriers, 8-step
    rings, two centre columns per statistics thread -- DESIGN.md section 5) instead of the classic streaming kernel: same strip
    walk and per-pixel forward arithmetic."""
    import e2e_slam_b200 as e2e
    from e2e_slam_b200 import ops
    from e2e_slam_b200.synthetic import make_pairs
    d = make_pairs(B, H, W, "tum", seed=H + W, rot_deg=3.0, trans=0.1)
    cu = {k: v.cuda() for k, v in d.items()}
    src0, tgt = cu["colors"][:, 0].permute(0, 3, 1, 2), cu["colors"][:, 1].permute(0, 3, 1, 2)

    def run(roles):
        monkeypatch.setenv("E2E_ROLES", "1" if roles else "0")
        depth = cu["depth"].clone().requires_grad_(True)
        src = src0.clone().requires_grad_(True)
        T = cu["T"].clone().requires_grad_(True)
        if disp:
            x = (1.0 / cu["depth"]).clone().requires_grad_(True)
            loss = ops.warp_photometric_loss_from_disparity(x, cu["inv_K"], cu["K"], T, src, tgt, torch.tensor(1.1, device="cuda"), "border", True)
            loss.backward()
            return loss.detach(), x.grad, src.grad, T.grad
        loss = e2e.warp_photometric_loss(depth, cu["inv_K"], cu["K"], T, src, tgt, "border", True)
        loss.backward()
        return loss.detach(), depth.grad, src.grad, T.grad

    l0, gd0, gs0, gt0 = run(False)
    l1, gd1, gs1, gt1 = run(True)
    # per-pixel values are the classic kernel's bits; the loss differs only by the order of the fp32 partial sums (the statistics
    # threads own two columns each), the gradients by the packed evaluation of the adjoint coefficients
    assert abs(float(l0) - float(l1)) <= 1e-6 * abs(float(l0))
    assert rel_max(gd1.cpu().numpy(), gd0.cpu().numpy()) <= 2e-6
    assert rel_max(gt1.cpu().numpy(), gt0.cpu().numpy()) <= 1e-5
    assert rel_max(gs1.cpu().numpy(), gs0.cpu().numpy()) <= 2e-6
