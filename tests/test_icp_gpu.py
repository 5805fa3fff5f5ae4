"""Point-to-plane ICP / GradICP odometry on the GPU against the numpy oracle
(oracle/icp_oracle.py: restated from the gradSLAM paper -- gradslam is not vendored by the reference, parity unpinned)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _scene(n_tgt=6000, n_src=1500, seed=0):
    """Target: points + normals on three mutually orthogonal, slightly rippled walls (constrains all six degrees of freedom).
    Source: a subset moved by the inverse of a known small rigid motion."""
    from oracle import icp_oracle as io
    rng = np.random.default_rng(seed)
    pts, nrm = [], []
    for axis in range(3):
        uv = rng.uniform(-1.0, 1.0, size=(n_tgt // 3, 2))
        h = 0.03 * np.sin(3.0 * uv[:, 0]) * np.cos(2.0 * uv[:, 1])
        gu = 0.09 * np.cos(3.0 * uv[:, 0]) * np.cos(2.0 * uv[:, 1])
        gv = -0.06 * np.sin(3.0 * uv[:, 0]) * np.sin(2.0 * uv[:, 1])
        p, n = np.zeros((len(uv), 3)), np.zeros((len(uv), 3))
        a, b = (axis + 1) % 3, (axis + 2) % 3
        p[:, a], p[:, b], p[:, axis] = uv[:, 0], uv[:, 1], -1.0 + h
        n[:, a], n[:, b], n[:, axis] = -gu, -gv, 1.0
        pts.append(p)
        nrm.append(n / np.linalg.norm(n, axis=1, keepdims=True))
    tgt, tn = np.concatenate(pts).astype(np.float32), np.concatenate(nrm).astype(np.float32)
    T_true = io.se3_exp(np.array([0.02, -0.015, 0.01, 0.01, -0.02, 0.015]))
    sel = rng.choice(len(tgt), n_src, replace=False)
    Ti = np.linalg.inv(T_true)
    src = (tgt[sel].astype(np.float64) @ Ti[:3, :3].T + Ti[:3, 3]).astype(np.float32)
    return src, tgt, tn, T_true


@pytest.mark.parametrize("grad_icp", [False, True])
def test_library_icp_matches_oracle_and_recovers_motion(grad_icp):
    from e2e_slam_b200 import odometry
    from oracle import icp_oracle as io
    src, tgt, tn, T_true = _scene()
    c = lambda a: torch.from_numpy(a).cuda()[None]
    eye = torch.eye(4, device="cuda")
    with torch.no_grad():
        if grad_icp:
            T, idx = odometry.point_to_plane_gradICP(c(src), c(tgt), c(tn), eye, numiters=20, nu=0.05)
            T_ref, idx_ref, errs = io.point_to_plane_gradicp(src, tgt, tn, np.eye(4), numiters=20, nu=0.05)
        else:
            T, idx = odometry.point_to_plane_ICP(c(src), c(tgt), c(tn), eye, numiters=20)
            T_ref, idx_ref, errs = io.point_to_plane_icp(src, tgt, tn, np.eye(4), numiters=20)
    T = T.cpu().numpy().astype(np.float64)
    assert errs[-1] < 1e-3 * errs[0]                                      # the oracle converged
    assert np.abs(T - T_ref).max() <= 2e-5                                # tolerance: fp32 normal equations vs float64
    assert np.abs(T - T_true).max() <= 1e-3                               # and it is the motion that was applied
    assert (idx.cpu().numpy() == idx_ref).mean() > 0.995                  # correspondences of the last iteration (near-ties may differ)


@pytest.mark.parametrize("grad_icp,thresh", [(True, None), (False, None), (True, 0.02), (False, 0.004)])
def test_library_reverse_sweep_matches_torch_autograd(grad_icp, thresh):
    """e2e_icp_backward (the reverse sweep of the device-side loop) against torch autograd through the same iteration written with
    torch ops (`*_torch`): gradients w.r.t. the source cloud, the target cloud, the target normals and the initial transform of a
    random linear functional of the recovered pose.  Both routes treat the correspondences as constants."""
    from e2e_slam_b200 import odometry
    src, tgt, tn, _ = _scene(n_tgt=3000, n_src=600, seed=3)
    # 2 mm of noise on the source cloud: without it the residuals vanish at convergence and the derivatives w.r.t. the target
    # normals are rounding noise (1e-8) on both routes
    src = (src + np.random.default_rng(7).normal(0.0, 2e-3, size=src.shape)).astype(np.float32)
    c = lambda a: torch.from_numpy(a).cuda()[None]
    g = torch.Generator(device="cuda").manual_seed(5)
    wgt = torch.randn(3, 4, generator=g, device="cuda")
    T0 = torch.eye(4, device="cuda")
    T0[:3, 3] = torch.tensor([0.004, -0.003, 0.002], device="cuda")
    kw = dict(numiters=6, dist_thresh=thresh)
    if grad_icp:
        kw.update(nu=0.05)
    out = []
    for route in ("lib", "torch"):
        leaves = [c(src).requires_grad_(True), c(tgt).requires_grad_(True), c(tn).requires_grad_(True), T0.clone().requires_grad_(True)]
        name = ("point_to_plane_gradICP" if grad_icp else "point_to_plane_ICP") + ("_torch" if route == "torch" else "")
        T, idx = getattr(odometry, name)(*leaves, **kw)
        (T[:3] * wgt).sum().backward()
        out.append((T.detach(), idx, [l.grad for l in leaves]))
    (T_l, idx_l, g_l), (T_t, idx_t, g_t) = out
    assert (T_l - T_t).abs().max() <= 2e-5
    assert (idx_l == idx_t).float().mean() > 0.99
    g_l[3], g_t[3] = g_l[3][:3], g_t[3][:3]      # the last row of a rigid transform is constant: the library returns 0 for it
    for name, a, b in zip(("src", "tgt", "normals", "T_init"), g_l, g_t):
        assert a is not None and bool(torch.isfinite(a).all())
        scale = float(b.abs().max())
        assert scale > 1e-6, (name, scale)
        err = float((a - b).abs().max()) / scale
        assert err <= 2e-3, (name, err)          # fp32 normal equations on both sides, float64 solve / exponential in ours


def test_torch_route_matches_library_and_is_differentiable():
    from e2e_slam_b200 import odometry
    src, tgt, tn, _ = _scene(n_tgt=3000, n_src=600, seed=3)
    c = lambda a: torch.from_numpy(a).cuda()[None]
    eye = torch.eye(4, device="cuda")
    with torch.no_grad():
        T_lib, _ = odometry.point_to_plane_gradICP(c(src), c(tgt), c(tn), eye, numiters=8, nu=0.05)
        T_icp, _ = odometry.point_to_plane_ICP(c(src), c(tgt), c(tn), eye, numiters=8)
    s = c(src).requires_grad_(True)
    T, _ = odometry.point_to_plane_gradICP(s, c(tgt), c(tn), eye, numiters=8, nu=0.05)
    assert (T.detach() - T_lib).abs().max() <= 2e-5
    T[:3, 3].sum().backward()
    g = s.grad
    assert g is not None and bool(torch.isfinite(g).all()) and float(g.abs().sum()) > 0
    s2 = c(src).requires_grad_(True)
    T2, _ = odometry.point_to_plane_ICP(s2, c(tgt), c(tn), eye, numiters=8)
    assert (T2.detach() - T_icp).abs().max() <= 2e-5
    st = c(src).requires_grad_(True)
    Tt, _ = odometry.point_to_plane_gradICP_torch(st, c(tgt), c(tn), eye, numiters=8, nu=0.05)
    assert (Tt.detach() - T_lib).abs().max() <= 2e-5
    # finite-difference check of d(translation sum)/d(one source coordinate) through the differentiable route
    with torch.no_grad():
        eps = 1e-3
        sp = c(src).clone(); sp[0, 5, 1] += eps
        sm = c(src).clone(); sm[0, 5, 1] -= eps
        fp = odometry.point_to_plane_gradICP(sp.requires_grad_(False), c(tgt), c(tn), eye, numiters=8, nu=0.05)[0][:3, 3].sum()
        fm = odometry.point_to_plane_gradICP(sm, c(tgt), c(tn), eye, numiters=8, nu=0.05)[0][:3, 3].sum()
    fd = float(fp - fm) / (2 * eps)
    assert abs(fd - float(g[0, 5, 1])) <= 0.2 * max(abs(fd), 1e-4) + 1e-4


def test_interface_errors():
    from e2e_slam_b200 import odometry
    a = torch.rand(1, 10, 3).cuda()
    with pytest.raises(ValueError):
        odometry.point_to_plane_ICP(a[0], a, a, torch.eye(4).cuda())
    with pytest.raises(ValueError):
        odometry.point_to_plane_ICP(a, a, a[:, :5], torch.eye(4).cuda())
    with pytest.raises(ValueError):
        odometry.point_to_plane_ICP(a, a, a, torch.eye(3).cuda())
    with pytest.raises(TypeError):
        odometry.point_to_plane_gradICP(a.double(), a, a, torch.eye(4).cuda())
    with pytest.raises(ValueError):
        odometry.se3_exp(torch.zeros(5).cuda())


@pytest.mark.parametrize("odom", ["icp", "gradicp"])
def test_pointfusion_tracks_a_sequence_without_ground_truth_poses(odom):
    """PointFusion(odom=...) as the reference constructs it (train_depth.py:111-116): only the first pose is given, every
    later frame is localised by ICP against the map.  The recovered trajectory must stay within 3 cm of the ground truth
    of the synthetic room over 13 cm of travel (1200 points per alignment, box room: point-to-plane is weakly
    constrained along the walls) and beat the constant-pose guess by a wide margin."""
    from e2e_slam_b200.slam import PointFusion, Pointclouds, RGBDImages
    from e2e_slam_b200.synthetic import room_sequence
    L, H, W = 6, 120, 160
    depth, rgb, K, poses = room_sequence(L, H, W, device="cuda")
    first = poses[:1].clone()
    rgbd = RGBDImages(rgb[None], depth[None, ..., None], K.view(1, 1, 4, 4), None)
    slam = PointFusion(odom=odom, dsratio=4, numiters=20, nu=0.05, device="cuda")
    pc, prev, out = Pointclouds(device="cuda"), None, []
    with torch.no_grad():
        for s in range(L):
            live = rgbd[:, s]
            if s == 0:
                live.poses = first[None]
            pc, pose = slam.step(pc, live, prev, inplace=True)
            prev = live
            out.append(pose[0, 0])
    out = torch.stack(out)
    err = (out[:, :3, 3] - poses[:, :3, 3]).norm(dim=1)
    drift_if_static = (poses[:, :3, 3] - poses[0, :3, 3]).norm(dim=1)
    assert float(err.max()) < 0.03, err
    assert float(err[-1]) < 0.3 * float(drift_if_static[-1])
    assert float((out[:, :3, :3] - poses[:, :3, :3]).abs().max()) < 1e-2


def test_gradicp_pose_is_differentiable_wrt_the_live_depth(monkeypatch):
    """GradICP's purpose (online_adaption.py:362-363 with odom = "gradicp"): the pose PointFusion recovers for a live frame carries
    a gradient back to that frame's depth map -- through the library's reverse sweep.  The same step with the torch-op iteration in
    place of the library call (torch autograd end to end) must give the same depth gradient.  (A finite difference of the fp32 pose
    is too noisy to check a derivative of 2e-3: measured +-3e-4 at a 1 mm step.)"""
    from e2e_slam_b200 import odometry
    from e2e_slam_b200.slam import PointFusion, Pointclouds, RGBDImages
    from e2e_slam_b200.synthetic import room_sequence
    L, H, W = 2, 120, 160
    depth, rgb, K, poses = room_sequence(L, H, W, device="cuda")
    slam = PointFusion(odom="gradicp", dsratio=4, numiters=6, nu=0.05, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(2)
    wgt = torch.randn(3, 4, generator=g, device="cuda")

    def depth_grad():
        rgbd0 = RGBDImages(rgb[None, :1], depth[None, :1, ..., None], K.view(1, 1, 4, 4), poses[None, :1])
        pc, _ = slam.step(Pointclouds(device="cuda"), rgbd0, None, inplace=False)
        d1 = depth[1].clone().requires_grad_(True)
        live = RGBDImages(rgb[None, 1:2], d1[None, None, ..., None], K.view(1, 1, 4, 4), None)
        pose = slam._localize(pc, live, rgbd0)[0, 0]
        (pose[:3] * wgt).sum().backward()
        return pose.detach(), d1.grad

    pose_l, g_l = depth_grad()
    monkeypatch.setattr(odometry, "point_to_plane_gradICP", odometry.point_to_plane_gradICP_torch)
    pose_t, g_t = depth_grad()
    assert (pose_l - pose_t).abs().max() <= 2e-5
    assert bool(torch.isfinite(g_l).all()) and float(g_t.abs().max()) > 0
    assert float((g_l - g_t).abs().max()) <= 2e-3 * float(g_t.abs().max())
