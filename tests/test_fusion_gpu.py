"""GPU parity tests of PointFusion (vertex/normal/confidence maps, association index map, merge + append)
and of the K = 1 nearest-neighbour kernels, against oracle/fusion_oracle.py (numpy; gradslam / chamferdist
semantics as frozen there -- parity unpinned, see its header).  Integer outputs must be BIT-EXACT."""
import math

import numpy as np
import pytest
import torch

from conftest import rel_max, same_values

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _room(L, H, W, holes=0.0, seed=0):
    from oracle import fusion_oracle as fo
    depth, rgb, K, poses = fo.synthetic_room_sequence(L, H, W, seed)
    if holes > 0:
        rng = np.random.default_rng(seed + 1)
        depth = depth * (rng.random(depth.shape) >= holes)
    return depth.astype(np.float32), rgb.astype(np.float32), K, poses


def _rgbd(depth, rgb, K, poses, sl=slice(None)):
    from e2e_slam_b200.slam import RGBDImages
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return RGBDImages(t(rgb[sl])[None], t(depth[sl])[None, ..., None], t(K).view(1, 1, 4, 4), t(poses[sl])[None])


def test_rgbd_maps_bit_exact():
    from e2e_slam_b200.slam import _frame_maps
    from oracle import fusion_oracle as fo
    depth, rgb, K, poses = _room(3, 60, 80, holes=0.15)
    for s in range(3):
        o = fo.rgbd_maps(depth[s], rgb[s], K, poses[s], 0.6)
        vg, ng, alpha, valid = _frame_maps(torch.from_numpy(depth[s]).cuda(), torch.from_numpy(K).cuda(),
                                           torch.from_numpy(poses[s]).cuda(), 0.6)
        assert same_values(vg.cpu().numpy(), o["vertex_g"]) == 0
        assert same_values(ng.cpu().numpy(), o["normal_g"]) == 0
        assert same_values(alpha.cpu().numpy(), o["alpha"]) == 0
        assert np.array_equal(valid.cpu().numpy().astype(bool), o["valid"])


@pytest.mark.parametrize("L,H,W,holes", [(8, 60, 80, 0.0), (6, 120, 160, 0.15), (3, 480, 640, 0.1)])
def test_fusion_sequence_bit_exact(L, H, W, holes):
    """Frame-by-frame PointFusion.step from an empty map (slam/custom_slam.py:26-34 drives it this way):
    index map, append slots and the whole map after every step must equal the oracle's bit for bit."""
    from e2e_slam_b200.slam import PointFusion, Pointclouds
    from oracle import fusion_oracle as fo
    depth, rgb, K, poses = _room(L, H, W, holes)
    rgbd = _rgbd(depth, rgb, K, poses)
    slam = PointFusion(odom="gt", dist_th=0.05, angle_th=20, sigma=0.6, device="cuda")
    orc = fo.PointFusionOracle(0.05, 20, 0.6)
    pc = Pointclouds(device="cuda")
    matched_total = 0
    with torch.no_grad():
        for s in range(L):
            pc, _ = slam.step(pc, rgbd[:, s], inplace=(s % 2 == 0))       # exercise both in-place and functional paths
            o = orc.step(depth[s], rgb[s], K, poses[s])
            index_map, slot = slam.last_association[0]
            assert np.array_equal(index_map.cpu().numpy(), o["index_map"]), f"index map differs at frame {s}"
            assert np.array_equal(slot.cpu().numpy(), o["append_slot"]), f"append order differs at frame {s}"
            matched_total += int((o["index_map"] >= 0).sum())
            assert pc.points_list[0].shape[0] == len(orc.points)
            assert same_values(pc.points_list[0].cpu().numpy(), orc.points) == 0
            assert same_values(pc.normals_list[0].cpu().numpy(), orc.normals) == 0
            assert same_values(pc.colors_list[0].cpu().numpy(), orc.colors) == 0
            assert same_values(pc.features_list[0][:, 0].cpu().numpy(), orc.ccount) == 0
    assert matched_total > 0.5 * (L - 1) * (depth[0] > 0).sum()      # the test really exercises association


def test_fusion_call_full_sequence_invariants():
    """Config C3 shape (60 frames, 480x640) through PointFusion.__call__: size-independent properties --
    every valid live pixel ends up in exactly one map point, so sum(ccount) == sum over frames of alpha over
    valid pixels, and N == number of appended pixels; confidence counts are positive."""
    from e2e_slam_b200.slam import PointFusion, _frame_maps
    from e2e_slam_b200.synthetic import room_sequence
    from e2e_slam_b200.slam import RGBDImages
    L, H, W = 60, 480, 640
    depth, rgb, K, poses = room_sequence(L, H, W, device="cuda")
    depth[:, 100:140, 200:260] = 0.0                                  # a hole, like TUM
    rgbd = RGBDImages(rgb[None], depth[None, ..., None], K.view(1, 1, 4, 4), poses[None])
    slam = PointFusion(odom="gt", device="cuda")
    with torch.no_grad():
        pc, out_poses = slam(rgbd)
        alpha_total = 0.0
        for s in range(L):
            _, _, alpha, valid = _frame_maps(depth[s].contiguous(), K, poses[s].contiguous(), 0.6)
            alpha_total += float((alpha.double() * valid.double()).sum())
    cc = pc.features_list[0][:, 0]
    assert torch.equal(out_poses, poses[None])
    assert float(cc.min()) > 0
    assert abs(float(cc.double().sum()) - alpha_total) <= 1e-5 * alpha_total
    n = pc.points_list[0].shape[0]
    assert (depth[0] > 0).sum() <= n < 0.25 * L * H * W                # map grows, but most pixels are merged
    assert bool(torch.isfinite(pc.points_list[0]).all())


def test_sequence_call_equals_step_by_step():
    """PointFusion.forward with known poses and no autograd runs the frame loop inside the library
    (e2e_fusion_sequence); the map must be, bit for bit, the one the per-frame step() calls build."""
    from e2e_slam_b200.slam import PointFusion, Pointclouds, RGBDImages
    from e2e_slam_b200.synthetic import room_sequence
    L, H, W = 12, 120, 160
    depth, rgb, K, poses = room_sequence(L, H, W, device="cuda")
    depth[:, 30:50, 60:90] = 0.0
    rgbd = RGBDImages(rgb[None], depth[None, ..., None], K.view(1, 1, 4, 4), poses[None])
    with torch.no_grad():
        fast, out_poses = PointFusion(odom="gt", device="cuda")(rgbd)
        slam = PointFusion(odom="gt", device="cuda")
        pc = Pointclouds(device="cuda")
        for s in range(L):
            pc, _ = slam.step(pc, rgbd[:, s], inplace=True)
    assert torch.equal(out_poses, poses[None])
    for a, b in ((fast.points_list, pc.points_list), (fast.normals_list, pc.normals_list), (fast.colors_list, pc.colors_list),
                 (fast.features_list, pc.features_list)):
        assert a[0].shape == b[0].shape and torch.equal(a[0], b[0])
    PointFusion(odom="gt", device="cuda").compact(fast)
    assert fast._maps[0].pts.shape[0] == pc.points_list[0].shape[0]
    assert torch.equal(fast.points_list[0], pc.points_list[0])


def _torch_fuse(depth, rgb, K, pose, old, index_map, append_slot, sigma=0.6):
    """float64 torch restatement of one fusion step given the (integer) association, for autograd."""
    H, W = depth.shape
    u = torch.arange(W, dtype=torch.float64)[None, :].expand(H, W)
    v = torch.arange(H, dtype=torch.float64)[:, None].expand(H, W)
    X = (u / K[0, 0] - K[0, 2] / K[0, 0]) * depth
    Y = (v / K[1, 1] - K[1, 2] / K[1, 1]) * depth
    m = (depth > 0).double()
    V = torch.stack([X, Y, depth], -1) * m[..., None]
    Vg = (V @ pose[:3, :3].t() + pose[:3, 3]) * m[..., None]
    alpha = torch.exp(-(V[..., 0] ** 2 + V[..., 1] ** 2) / (2 * sigma ** 2))
    pts, col, cc = old
    h, w = torch.nonzero(index_map >= 0, as_tuple=True)
    n = index_map[h, w]
    c, a = cc[n][:, None], alpha[h, w][:, None]
    pts, col, cc = pts.clone(), col.clone(), cc.clone()
    pts[n] = (c * pts[n] + a * Vg[h, w]) / (c + a)
    col[n] = (c * col[n] + a * rgb[h, w]) / (c + a)
    cc[n] = (c + a)[:, 0]
    new = append_slot >= 0
    return torch.cat([pts, Vg[new]]), torch.cat([col, rgb[new]]), torch.cat([cc, alpha[new]])


def test_fusion_gradients_reach_depth_and_colour():
    """gradient_experiments.py:120-160 / image_recover_slam: fuse two frames, the first detached; a loss on
    the fused points, colours and confidences must back-propagate to the last frame's depth and rgb.  Also
    the empty-map step of online_adaption.py:457-471.  Truth: autograd on a float64 torch restatement that is
    handed the oracle's integer association."""
    from e2e_slam_b200.slam import PointFusion, image_recover_slam
    from oracle import fusion_oracle as fo
    L, H, W = 2, 48, 64
    depth, rgb, K, poses = _room(L, H, W, holes=0.1)
    rgbd = _rgbd(depth, rgb, K, poses)
    rgbd.depth_image.requires_grad_(True)
    rgbd.rgb_image.requires_grad_(True)
    slam = PointFusion(odom="gt", device="cuda")
    pc = image_recover_slam(rgbd, slam, torch.device("cuda"))
    P, C, F = pc.points_list[0], pc.colors_list[0], pc.features_list[0][:, 0]
    torch.manual_seed(0)
    wp, wc, wf = torch.randn(P.shape), torch.randn(C.shape), torch.randn(F.shape)
    ((P * wp.cuda()).sum() + (C * wc.cuda()).sum() + (F * wf.cuda()).sum()).backward()
    g_depth, g_rgb = rgbd.depth_image.grad[0, :, :, :, 0].cpu().numpy(), rgbd.rgb_image.grad[0].cpu().numpy()
    assert np.all(g_depth[0] == 0) and np.all(g_rgb[0] == 0)         # custom_slam.py:27-28: only the last frame
    # truth
    orc = fo.PointFusionOracle()
    orc.step(depth[0], rgb[0], K, poses[0])
    old = tuple(torch.from_numpy(a).double() for a in (orc.points, orc.colors, orc.ccount))
    o = orc.step(depth[1], rgb[1], K, poses[1])
    d1 = torch.from_numpy(depth[1]).double().requires_grad_(True)
    c1 = torch.from_numpy(rgb[1]).double().requires_grad_(True)
    p64, c64, f64 = _torch_fuse(d1, c1, torch.from_numpy(K).double(), torch.from_numpy(poses[1]).double(), old,
                                torch.from_numpy(o["index_map"]), torch.from_numpy(o["append_slot"]))
    ((p64 * wp.double()).sum() + (c64 * wc.double()).sum() + (f64 * wf.double()).sum()).backward()
    assert rel_max(g_depth[1], d1.grad.numpy()) <= RTOL
    assert rel_max(g_rgb[1], c1.grad.numpy()) <= RTOL


def test_fusion_gradients_through_the_old_map():
    """PointFusion.__call__ with every frame requiring grad (train_depth.py:378-384 + knn_points, :683):
    the gradient must also flow through the merged map points into EARLIER frames' depth."""
    from e2e_slam_b200.slam import PointFusion
    from oracle import fusion_oracle as fo
    L, H, W = 3, 40, 56
    depth, rgb, K, poses = _room(L, H, W)
    rgbd = _rgbd(depth, rgb, K, poses)
    rgbd.depth_image.requires_grad_(True)
    slam = PointFusion(odom="gt", device="cuda")
    pc, _ = slam(rgbd)
    P = pc.points_list[0]
    torch.manual_seed(1)
    wp = torch.randn(P.shape)
    (P * wp.cuda()).sum().backward()
    g = rgbd.depth_image.grad[0, :, :, :, 0].cpu().numpy()
    orc = fo.PointFusionOracle()
    ds = [torch.from_numpy(depth[s]).double().requires_grad_(True) for s in range(L)]
    state = (torch.zeros(0, 3, dtype=torch.float64), torch.zeros(0, 3, dtype=torch.float64), torch.zeros(0, dtype=torch.float64))
    for s in range(L):
        o = orc.step(depth[s], rgb[s], K, poses[s])
        state = _torch_fuse(ds[s], torch.from_numpy(rgb[s]).double(), torch.from_numpy(K).double(),
                            torch.from_numpy(poses[s]).double(), state, torch.from_numpy(o["index_map"]),
                            torch.from_numpy(o["append_slot"]))
    (state[0] * wp.double()).sum().backward()
    for s in range(L):
        assert rel_max(g[s], ds[s].grad.numpy()) <= RTOL, s


def test_pointfusion_interface_errors():
    from e2e_slam_b200.slam import PointFusion, Pointclouds, RGBDImages
    depth, rgb, K, poses = _room(2, 16, 20)
    rgbd = _rgbd(depth, rgb, K, poses)
    with pytest.raises(ValueError):
        PointFusion(odom="slam")
    with pytest.raises(TypeError):
        PointFusion(odom="gt").step([], rgbd[:, 0])
    with pytest.raises(ValueError):
        PointFusion(odom="gt").step(Pointclouds(device="cuda"), rgbd)             # sequence length 2
    with pytest.raises(ValueError):                                               # ICP odometry against an empty map
        PointFusion(odom="gradicp", device="cuda").step(Pointclouds(device="cuda"), rgbd[:, 1], rgbd[:, 0])
    pc, p = PointFusion(odom="gradicp", device="cuda").step(Pointclouds(device="cuda"), rgbd[:, 0], prev_frame=None)
    assert pc.has_points and p.shape == (1, 1, 4, 4)
    assert rgbd[:, 1].shape == (1, 1, 16, 20) and rgbd.shape == (1, 2, 16, 20)
    with pytest.raises(ValueError):
        RGBDImages(rgbd.rgb_image, rgbd.depth_image, rgbd.intrinsics[:, :, :3])


# ---- K = 1 nearest neighbour ------------------------------------------------------------------------
def test_knn_bit_exact_and_gradients():
    from e2e_slam_b200.losses import color_points_loss, knn_points, knn_points_loss, point_supervision_loss
    from oracle import fusion_oracle as fo
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(0)
    q = rng.normal(size=(5003, 3)).astype(np.float32)
    r = rng.normal(size=(7001, 3)).astype(np.float32)
    r[100] = r[50]; r[6000] = r[50]                                   # exact duplicates: lowest index must win
    q[7] = r[50]
    d2, idx = fo.knn1(q, r)
    out = knn_points(torch.from_numpy(q).cuda()[None], torch.from_numpy(r).cuda()[None])
    assert out.dists.shape == (1, 5003, 1) and out.idx.shape == (1, 5003, 1) and out.idx.dtype == torch.int64
    assert np.array_equal(out.idx[0, :, 0].cpu().numpy(), idx)
    assert same_values(out.dists[0, :, 0].cpu().numpy(), d2) == 0
    assert idx[7] == 50
    tree_d, tree_i = cKDTree(r.astype(np.float64)).query(q.astype(np.float64))
    agree = (tree_i == idx) | (np.abs(tree_d ** 2 - d2) <= 1e-6 * np.maximum(d2, 1e-12))     # fp32 near-ties may differ
    assert agree.all()
    # loss + gradients (loss/losses.py:39-63): mean of squared distances, autograd to both clouds
    qt, rt = torch.from_numpy(q).cuda().requires_grad_(True), torch.from_numpy(r).cuda().requires_grad_(True)
    loss, indexes = knn_points_loss(rt[None], qt[None])               # (gt, noisy) argument order of the reference
    loss.backward()
    q64, r64 = torch.from_numpy(q).double().requires_grad_(True), torch.from_numpy(r).double().requires_grad_(True)
    ((q64 - r64[torch.from_numpy(idx)]) ** 2).sum(1).mean().backward()
    assert abs(float(loss) - float(d2.astype(np.float64).mean())) <= RTOL * float(d2.mean())
    assert np.array_equal(indexes[0].cpu().numpy(), idx)
    assert rel_max(qt.grad.cpu().numpy(), q64.grad.numpy()) <= RTOL
    assert rel_max(rt.grad.cpu().numpy(), r64.grad.numpy()) <= RTOL
    # colour loss (losses.py:65-82)
    cq, cr = torch.rand(1, 5003, 3).cuda(), torch.rand(1, 7001, 3).cuda()
    cl = color_points_loss(cr, cq, indexes)
    ref = (cq[0].cpu() - cr[0].cpu()[torch.from_numpy(idx)]).abs().mean()
    assert abs(float(cl) - float(ref)) <= 1e-6
    # fused transform (online_adaption.py:638-645)
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = np.array([[0.9998, -0.0175, 0.0], [0.0175, 0.9998, 0.0], [0, 0, 1]], np.float32)
    T[:3, 3] = [0.01, -0.02, 0.03]
    d2t, idxt = fo.knn1(fo.transform_pointcloud(q, T), r)
    qt2 = torch.from_numpy(q).cuda().requires_grad_(True)
    l3 = point_supervision_loss(qt2, torch.from_numpy(T).cuda(), torch.from_numpy(r).cuda())
    l3.backward()
    assert abs(float(l3) - float(d2t.astype(np.float64).mean())) <= RTOL * float(d2t.mean())
    q64 = torch.from_numpy(q).double().requires_grad_(True)
    T64 = torch.from_numpy(T).double()
    (((q64 @ T64[:3, :3].t() + T64[:3, 3]) - torch.from_numpy(r).double()[torch.from_numpy(idxt)]) ** 2).sum(1).mean().backward()
    assert rel_max(qt2.grad.cpu().numpy(), q64.grad.numpy()) <= RTOL


def test_knn_errors_and_ragged_sizes():
    from e2e_slam_b200.losses import knn_points_loss
    from oracle import fusion_oracle as fo
    a, b = torch.rand(1, 10, 3).cuda(), torch.rand(2, 10, 3).cuda()
    with pytest.raises(ValueError):
        knn_points_loss(a, b)
    with pytest.raises(ValueError):
        knn_points_loss(a, torch.rand(1, 10, 2).cuda())
    for p1, p2 in ((1, 1), (3, 2049), (1025, 5), (4097, 2048)):       # around the tile / thread-block edges
        q, r = np.random.default_rng(p1).random((p1, 3), np.float32), np.random.default_rng(p2 + 1).random((p2, 3), np.float32)
        l, i = knn_points_loss(torch.from_numpy(r).cuda()[None], torch.from_numpy(q).cuda()[None])
        d2, idx = fo.knn1(q, r)
        assert np.array_equal(i[0].cpu().numpy(), idx)


def test_point_supervision_full_size():
    """Config C2 shape: 307 200 query points (one live frame) against a 200 k-point map; idx spot-checked
    with a KD-tree (float64) and the loss checked against it."""
    from e2e_slam_b200.losses import knn_points
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(3)
    q = (rng.random((307200, 3)) * 4).astype(np.float32)
    r = (rng.random((200000, 3)) * 4).astype(np.float32)
    out = knn_points(torch.from_numpy(q).cuda()[None], torch.from_numpy(r).cuda()[None])
    d, i = cKDTree(r.astype(np.float64)).query(q.astype(np.float64))
    ours_i, ours_d = out.idx[0, :, 0].cpu().numpy(), out.dists[0, :, 0].cpu().numpy()
    assert (ours_i == i).mean() > 0.9999
    assert np.abs(ours_d - d ** 2).max() <= 1e-5 * (d ** 2).max()


@pytest.mark.parametrize("H,W", [(16, 20), (120, 160), (270, 480), (1080, 1920)])
def test_sequence_entry_continues_an_existing_map(H, W):
    """e2e_fusion_sequence on frames 0..5 in one call == frames 0..2, then a second call that starts from that map
    (the kernel's prologue converts the caller's [n,3] arrays to its working records): every array bit for bit.
    Sizes: a frame smaller than one CTA's share, the usual quarter frame, and one with several sub-blocks per CTA."""
    import ctypes
    from e2e_slam_b200._lib import check, lib, ptr, stream_ptr
    from e2e_slam_b200.synthetic import room_sequence
    L = 6
    depth, rgb, K, poses = room_sequence(L, H, W, device="cuda")
    depth[:, H // 4:H // 3, W // 3:W // 2] = 0.0

    def run(frames, state=None):
        cap = L * H * W
        z = dict(dtype=torch.float32, device="cuda")
        if state is None:
            pts, nrm, col, cc = torch.zeros(cap, 3, **z), torch.zeros(cap, 3, **z), torch.zeros(cap, 3, **z), torch.zeros(cap, **z)
            n, upper = torch.zeros(2, dtype=torch.int64, device="cuda"), 0
        else:
            pts, nrm, col, cc, n = state
            upper = int(n[0])
        nws = lib().e2e_fusion_sequence_workspace_bytes(H, W, cap)
        ws = torch.empty(nws, dtype=torch.uint8, device="cuda")
        d, c, p = depth[frames].contiguous(), rgb[frames].contiguous(), poses[frames].contiguous()
        check(lib().e2e_fusion_sequence(ptr(d), ptr(c), ptr(K), ptr(p), d.shape[0], H, W, ctypes.c_float(0.6), ctypes.c_float(0.05),
                                        ctypes.c_float(math.cos(math.radians(20))), ptr(pts), ptr(nrm), ptr(col), ptr(cc), ptr(n),
                                        upper, cap, ptr(ws), nws, stream_ptr()), "e2e_fusion_sequence")
        torch.cuda.synchronize()
        return pts, nrm, col, cc, n

    whole = run(slice(0, 6))
    part = run(slice(3, 6), run(slice(0, 3)))
    n = int(whole[4][0])
    assert n == int(part[4][0]) and n > (depth[0] > 0).sum()
    for a, b in zip(whole[:4], part[:4]):
        assert torch.equal(a[:n], b[:n])


def _knn_both(q, r, T=None):
    """(dist2, idx) from the brute-force kernel and from the grid kernel, through the C ABI."""
    import ctypes
    from e2e_slam_b200._lib import check, lib, ptr, stream_ptr
    q, r = q.contiguous(), r.contiguous()
    P1, P2 = q.shape[0], r.shape[0]
    out = []
    for grid in (False, True):
        d2 = torch.empty(P1, dtype=torch.float32, device="cuda")
        idx = torch.empty(P1, dtype=torch.int64, device="cuda")
        if grid:
            nws = lib().e2e_knn1_grid_workspace_bytes(P2)
            ws = torch.empty(nws, dtype=torch.uint8, device="cuda")
            check(lib().e2e_knn1_grid_fwd(ptr(q), ptr(T), ptr(r), P1, P2, ptr(d2), ptr(idx), ptr(ws), nws, stream_ptr()), "grid")
        else:
            check(lib().e2e_knn1_fwd(ptr(q), ptr(T), ptr(r), P1, P2, ptr(d2), ptr(idx), stream_ptr()), "brute")
        out.append((d2, idx))
    return out


@pytest.mark.parametrize("case", ["uniform", "surface", "clustered", "duplicates", "far_queries", "all_far", "off_surface", "line", "single",
                                  "nonfinite", "slanted_far", "box_edges", "plane", "huge_extent"])
def test_grid_knn_equals_brute_force(case):
    """The uniform-grid kernel must return the brute-force kernel's answer bit for bit (same distance arithmetic, lowest
    index among exact ties) whatever the shape of the clouds."""
    g = torch.Generator(device="cuda").manual_seed(11)
    rnd = lambda *s: torch.rand(*s, generator=g, device="cuda")
    T = None
    if case == "uniform":
        q, r = rnd(20011, 3) * 4, rnd(50021, 3) * 4
    elif case == "surface":                     # a depth-map like sheet: the real use (live frame vs. map), with a fused transform
        uv = rnd(60000, 2) * 4 - 2
        r = torch.stack([uv[:, 0], uv[:, 1], 2.5 + 0.3 * torch.sin(2 * uv[:, 0]) * torch.cos(uv[:, 1])], 1) + 0.002 * rnd(60000, 3)
        q = r[torch.randperm(60000, device="cuda", generator=g)[:15000]] + 0.01 * (rnd(15000, 3) - 0.5)
        T = torch.eye(4, device="cuda"); T[:3, 3] = torch.tensor([0.01, -0.02, 0.015], device="cuda"); T[0, 1], T[1, 0] = -0.01, 0.01
    elif case == "clustered":                   # a few dense blobs and a lot of empty space
        c = rnd(7, 3) * 50
        r = (c[torch.randint(0, 7, (40000,), device="cuda", generator=g)] + 0.05 * (rnd(40000, 3) - 0.5))
        q = rnd(9000, 3) * 50
    elif case == "duplicates":                  # many exact ties: the lowest index must win
        base = rnd(300, 3)
        r = base[torch.randint(0, 300, (20000,), device="cuda", generator=g)]
        q = torch.cat([base, rnd(3000, 3)])
    elif case == "far_queries":                 # queries far outside the reference box, on every side
        r = rnd(30000, 3)
        q = torch.cat([rnd(2000, 3) * 200 - 100, rnd(500, 3)])
    elif case == "all_far":                     # every query far outside the reference box (a wrong pose): must stay cheap
        r = rnd(30000, 3)
        q = rnd(20000, 3) + torch.tensor([40.0, -25.0, 60.0], device="cuda")
    elif case == "off_surface":                 # queries several cell sizes off a dense sheet (a noisy depth prediction)
        uv = rnd(200000, 2) * 2 - 1
        r = torch.stack([uv[:, 0], uv[:, 1], 2.0 + 0.2 * torch.sin(3 * uv[:, 0])], 1)
        q = r[torch.randperm(200000, device="cuda", generator=g)[:30000]].clone()
        q[:, 2] += 0.1 * torch.randn(30000, generator=g, device="cuda")
    elif case == "line":                        # degenerate extent on two axes
        r = torch.zeros(25000, 3, device="cuda"); r[:, 0] = rnd(25000) * 10
        q = torch.zeros(4000, 3, device="cuda"); q[:, 0] = rnd(4000) * 12 - 1; q[:, 1] = 0.01 * rnd(4000)
    elif case == "single":
        r, q = rnd(1, 3), rnd(100, 3) * 5
    elif case == "plane":                       # exactly flat: one cell along z, the coarse-major ids are padded to whole coarse cells
        r = torch.cat([rnd(90000, 2) * 6, torch.full((90000, 1), 1.25, device="cuda")], 1)
        q = torch.cat([rnd(8000, 2) * 7 - 0.5, 1.25 + 0.3 * (rnd(8000, 1) - 0.5)], 1)
    elif case == "huge_extent":                 # two far-apart blobs: the cell size is enlarged until the padded grid fits
        r = torch.cat([rnd(40000, 3) * 0.5, rnd(40000, 3) * 0.5 + 5000.0])
        q = torch.cat([rnd(3000, 3) * 0.6, rnd(3000, 3) * 0.6 + 5000.0, rnd(500, 3) * 5000.0])
    elif case == "slanted_far":                 # a steep, thick sheet and queries 0-60 cm off it: long ring walks, tight coarse boxes
        uv = rnd(300000, 2) * 4 - 2
        r = torch.stack([uv[:, 0], uv[:, 1], 3.0 + 0.9 * uv[:, 0] - 0.6 * uv[:, 1]], 1) + 0.01 * (rnd(300000, 3) - 0.5)
        q = r[torch.randperm(300000, device="cuda", generator=g)[:20000]].clone()
        q += 0.6 * (rnd(20000, 1) ** 2) * torch.nn.functional.normalize(torch.randn(20000, 3, generator=g, device="cuda"), dim=1)
    elif case == "box_edges":                   # queries on and just outside the faces, edges and corners of the reference box
        r = rnd(80000, 3)
        side = torch.randint(0, 3, (12000, 3), device="cuda", generator=g).float() * 0.5      # 0, 0.5 or 1 per axis
        q = side + 0.03 * (rnd(12000, 3) - 0.5)
        q[::7] = side[::7]                                                                     # exactly on the faces
    else:                                       # NaN / inf on either side never match; all-NaN queries answer (inf, 0) like brute force
        r = rnd(5000, 3); r[17] = float("nan"); r[99, 1] = float("inf")
        q = rnd(600, 3); q[3, 2] = float("nan"); q[40] = float("inf")
    (d_b, i_b), (d_g, i_g) = _knn_both(q, r, T)
    assert torch.equal(i_b, i_g), f"indices differ at {int((i_b != i_g).sum())} of {q.shape[0]} queries"
    assert torch.equal(d_b.view(torch.int32), d_g.view(torch.int32))


def test_grid_knn_full_size_point_supervision():
    """Config C2 / online_adaption shape: one live frame (307 200 points, moved by a small transform) against a 2 M-point
    map.  Brute force would be 6e11 pairs; the grid answer is checked against a float64 KD-tree."""
    from scipy.spatial import cKDTree
    from e2e_slam_b200 import losses
    g = torch.Generator(device="cuda").manual_seed(5)
    uv = torch.rand(2_000_000, 2, generator=g, device="cuda") * 8 - 4
    r = torch.stack([uv[:, 0], uv[:, 1], 3.0 + 0.4 * torch.sin(uv[:, 0]) * torch.cos(1.3 * uv[:, 1])], 1)
    sel = torch.randperm(2_000_000, device="cuda", generator=g)[:307_200]
    q = (r[sel] + 0.01 * (torch.rand(307_200, 3, generator=g, device="cuda") - 0.5)).requires_grad_(True)
    T = torch.eye(4, device="cuda"); T[:3, 3] = torch.tensor([0.02, -0.01, 0.03], device="cuda")
    assert losses._use_grid(q.shape[0], r.shape[0])
    loss = losses.point_supervision_loss(q, T, r)
    loss.backward()
    qt = (q.detach() @ T[:3, :3].t() + T[:3, 3]).cpu().double().numpy()
    d, i = cKDTree(r.cpu().double().numpy()).query(qt)
    assert abs(float(loss) - float((d ** 2).mean())) <= 1e-5 * float((d ** 2).mean())
    assert bool(torch.isfinite(q.grad).all()) and float(q.grad.abs().sum()) > 0


def test_grid_cache_follows_the_reference_cloud(monkeypatch):
    """The grid over the reference cloud is reused while the cloud is unchanged and rebuilt after an in-place write or for
    another tensor (a stale grid would silently return wrong neighbours)."""
    from e2e_slam_b200 import losses
    monkeypatch.setattr(losses, "KNN_MODE", "grid")
    g = torch.Generator(device="cuda").manual_seed(2)
    r = torch.rand(5000, 3, generator=g, device="cuda")
    q = torch.rand(700, 3, generator=g, device="cuda")

    def brute(ref):
        d = ((q[:, None, :] - ref[None, :, :]) ** 2).sum(-1)
        return d.argmin(1)

    i1 = losses.knn_points(q[None], r[None]).idx[0, :, 0]
    ws1 = losses._GRID_CACHE["last"][2]
    i2 = losses.knn_points(q[None], r[None].detach()).idx[0, :, 0]            # a new tensor object over the same storage: reused
    assert losses._GRID_CACHE["last"][2] is ws1 and torch.equal(i1, i2) and torch.equal(i1, brute(r))
    r[:2500] += 0.37                                                          # in-place change: the version counter moves
    i3 = losses.knn_points(q[None], r[None]).idx[0, :, 0]
    assert losses._GRID_CACHE["last"][2] is not ws1 and torch.equal(i3, brute(r))
    r2 = torch.rand(5000, 3, generator=g, device="cuda")                      # another cloud of the same shape
    i4 = losses.knn_points(q[None], r2[None]).idx[0, :, 0]
    assert torch.equal(i4, brute(r2))
