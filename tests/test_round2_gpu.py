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
