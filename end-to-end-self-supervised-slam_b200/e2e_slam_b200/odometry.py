"""Point-to-plane ICP / GradICP odometry with gradslam's names and signatures (gradslam.odometry.icputils:
`point_to_plane_ICP`, `point_to_plane_gradICP`, `gauss_newton_solve`, `solve_linear_system`; gradslam.geometry.se3utils:
`se3_exp`) -- what PointFusion runs when `odom` is "icp" / "gradicp", the reference's shipped default
(configs/config.yaml:30-34; train_depth.py:111-116, 378-381; online_adaption.py:362-363).

One route, with or without autograd (conventions frozen in oracle/icp_oracle.py; gradslam itself is not vendored by the reference):
the whole iteration loop runs inside the library (e2e_icp_point_to_plane: exact grid kNN (target gridded once per call), fused
Jacobian + normal equations, device-side 6x6 solve and se3 exponential, no host synchronisation).  With autograd (GradICP's
purpose: poses differentiable w.r.t. the live depth) the same loop also records its history (e2e_icp_point_to_plane_saved) and
`backward()` is the library's reverse sweep (e2e_icp_backward: four launches per iteration; the correspondences are constants
of the derivative, as in gradslam).  The iteration written with torch ops around the kNN kernel (`*_torch`) is kept as the
autograd reference the tests compare against.
There is no CPU route: tensors must be CUDA fp32.
"""
import ctypes

import torch

from ._lib import check, f32, lib, ptr, stream_ptr
from .losses import knn_points


def se3_exp(xi):
    """xi = [v, omega] (6,) -> 4x4.  R = I + sin t/t K + (1-cos t)/t^2 K^2, translation V v with
    V = I + (1-cos t)/t^2 K + (t-sin t)/t^3 K^2; first-order forms when t^2 < 1e-12."""
    if xi.shape != (6,):
        raise ValueError(f"xi must have shape (6,), got {tuple(xi.shape)}")
    v, w = xi[:3], xi[3:]
    z = torch.zeros((), dtype=xi.dtype, device=xi.device)
    K = torch.stack([torch.stack([z, -w[2], w[1]]), torch.stack([w[2], z, -w[0]]), torch.stack([-w[1], w[0], z])])
    t2 = (w * w).sum()
    eye = torch.eye(3, dtype=xi.dtype, device=xi.device)
    # branch-free (no host read of t2: this runs every iteration of the autograd ICP route): below t^2 = 1e-12 the first-order
    # forms R = I + K, V = I + K / 2; the large-angle coefficients are evaluated on a safe argument there
    small = t2 < 1e-12
    one = torch.ones((), dtype=xi.dtype, device=xi.device)
    t2s = torch.where(small, one, t2)
    t = torch.sqrt(t2s)
    K2 = K @ K
    a = torch.where(small, one, torch.sin(t) / t)
    b = (1 - torch.cos(t)) / t2s
    c = torch.where(small, torch.zeros_like(one), (t - torch.sin(t)) / (t2s * t))
    R = eye + a * K + torch.where(small, torch.zeros_like(one), b) * K2
    V = eye + torch.where(small, 0.5 * one, b) * K + c * K2
    top = torch.cat([R, (V @ v).unsqueeze(1)], 1)
    bottom = torch.tensor([[0.0, 0.0, 0.0, 1.0]], dtype=xi.dtype, device=xi.device)
    return torch.cat([top, bottom], 0)


def solve_linear_system(A, b, damp=1e-8):
    """(A^T A + damp I)^-1 A^T b for A (N,6), b (N,1) -> (6,1)."""
    if A.dim() != 2 or A.shape[1] != 6 or b.shape != (A.shape[0], 1):
        raise ValueError(f"expected A (N,6) and b (N,1), got {tuple(A.shape)} / {tuple(b.shape)}")
    At = A.t().double()
    M = At @ A.double() + torch.as_tensor(damp, dtype=torch.float64, device=A.device) * torch.eye(6, dtype=torch.float64, device=A.device)
    return torch.linalg.solve(M, At @ b.double()).to(A.dtype)


def _check_clouds(src_pc, tgt_pc, tgt_normals):
    for name, t in (("src_pc", src_pc), ("tgt_pc", tgt_pc), ("tgt_normals", tgt_normals)):
        if not torch.is_tensor(t):
            raise TypeError(f"Expected {name} to be of type torch.Tensor. Got {type(t)}.")
        if t.dim() != 3 or t.shape[0] != 1 or t.shape[2] != 3:
            raise ValueError(f"{name} should have shape (1, N, 3). Got {tuple(t.shape)}.")
        f32(t, name)
    if tgt_pc.shape != tgt_normals.shape:
        raise ValueError(f"tgt_pc and tgt_normals should have the same shape ({tuple(tgt_pc.shape)} != {tuple(tgt_normals.shape)})")


def gauss_newton_solve(src_pc, tgt_pc, tgt_normals, dist_thresh=None):
    """Linearised point-to-plane system of the current alignment: A (N',6) = [n, s x n], b (N',1) = n . (d - s) for every
    source point s with nearest target point d (normal n); pairs with squared distance >= dist_thresh are dropped.
    Returns (A, b, chamfer_indices)."""
    _check_clouds(src_pc, tgt_pc, tgt_normals)
    nn = knn_points(src_pc.detach().contiguous(), tgt_pc.detach().contiguous())
    d2, idx = nn.dists[0, :, 0], nn.idx[0, :, 0]
    s, d, n = src_pc[0], tgt_pc[0][idx], tgt_normals[0][idx]
    if dist_thresh is not None:
        keep = d2 < dist_thresh
        s, d, n = s[keep], d[keep], n[keep]
    A = torch.cat([n, torch.cross(s, n, dim=1)], 1)
    b = (n * (d - s)).sum(1, keepdim=True)
    return A, b, idx


def _apply(T, pts):
    return pts @ T[:3, :3].t() + T[:3, 3]


def _library_icp(src_pc, tgt_pc, tgt_normals, initial_transform, numiters, damp, dist_thresh, grad_icp, lambda_max, B, B2, nu):
    src, tgt, nrm = src_pc[0].contiguous(), tgt_pc[0].contiguous(), tgt_normals[0].contiguous()
    T0 = f32(initial_transform, "initial_transform").contiguous()
    dev = src.device
    N, M = src.shape[0], tgt.shape[0]
    T_out = torch.empty(4, 4, dtype=torch.float32, device=dev)
    idx = torch.zeros(N, dtype=torch.int64, device=dev) if int(numiters) <= 0 else torch.empty(N, dtype=torch.int64, device=dev)
    nws = lib().e2e_icp_workspace_bytes(N, M)
    ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib().e2e_icp_point_to_plane(ptr(src), N, ptr(tgt), ptr(nrm), M, ptr(T0), int(numiters), ctypes.c_float(damp),
                                           ctypes.c_float(-1.0 if dist_thresh is None else dist_thresh), int(grad_icp),
                                           ctypes.c_float(lambda_max), ctypes.c_float(B), ctypes.c_float(B2), ctypes.c_float(nu),
                                           ptr(T_out), ptr(idx), None, ptr(ws), nws, stream_ptr()), "e2e_icp_point_to_plane")
    return T_out, idx


def _needs_grad(*ts):
    return torch.is_grad_enabled() and any(t.requires_grad for t in ts)


class _ICPFunction(torch.autograd.Function):
    """The library's iteration loop as a differentiable operation: (src [N,3], tgt [M,3], normals [M,3], T0 [4,4]) -> (T [4,4], idx [N])."""

    @staticmethod
    def forward(ctx, src, tgt, nrm, T0, numiters, damp, dist_thresh, grad_icp, lambda_max, B, B2, nu):
        src, tgt, nrm, T0 = src.contiguous(), tgt.contiguous(), nrm.contiguous(), f32(T0, "initial_transform").contiguous()
        dev = src.device
        N, M = src.shape[0], tgt.shape[0]
        numiters = int(numiters)
        T_out = torch.empty(4, 4, dtype=torch.float32, device=dev)
        idx = torch.zeros(N, dtype=torch.int64, device=dev) if numiters <= 0 else torch.empty(N, dtype=torch.int64, device=dev)
        nws, nh = lib().e2e_icp_workspace_bytes(N, M), lib().e2e_icp_history_bytes(N, numiters)
        ws = torch.empty(nws, dtype=torch.uint8, device=dev)
        hist = torch.empty(nh, dtype=torch.uint8, device=dev)
        thr = -1.0 if dist_thresh is None else float(dist_thresh)
        with torch.cuda.device(dev):
            check(lib().e2e_icp_point_to_plane_saved(ptr(src), N, ptr(tgt), ptr(nrm), M, ptr(T0), numiters, ctypes.c_float(damp),
                                                     ctypes.c_float(thr), int(grad_icp), ctypes.c_float(lambda_max), ctypes.c_float(B),
                                                     ctypes.c_float(B2), ctypes.c_float(nu), ptr(T_out), ptr(idx), ptr(ws), nws,
                                                     ptr(hist), nh, stream_ptr()), "e2e_icp_point_to_plane_saved")
        ctx.save_for_backward(src, tgt, nrm, T0, hist)
        ctx.params = (numiters, float(damp), thr, int(grad_icp), float(lambda_max), float(B), float(B2), float(nu))
        ctx.mark_non_differentiable(idx)
        return T_out, idx

    @staticmethod
    def backward(ctx, gT, _gidx):
        src, tgt, nrm, T0, hist = ctx.saved_tensors
        numiters, damp, thr, grad_icp, lambda_max, B, B2, nu = ctx.params
        dev = src.device
        N, M = src.shape[0], tgt.shape[0]
        need = ctx.needs_input_grad
        gT = f32(gT, "grad_T").contiguous()
        g_src = torch.empty_like(src)
        g_tgt = torch.zeros_like(tgt) if need[1] else None      # accumulated with atomics
        g_nrm = torch.zeros_like(nrm) if need[2] else None
        g_T0 = torch.empty(4, 4, dtype=torch.float32, device=dev) if need[3] else None
        nws = lib().e2e_icp_backward_workspace_bytes(N)
        ws = torch.empty(nws, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib().e2e_icp_backward(ptr(src), N, ptr(tgt), ptr(nrm), M, ptr(T0), numiters, ctypes.c_float(damp), ctypes.c_float(thr),
                                         grad_icp, ctypes.c_float(lambda_max), ctypes.c_float(B), ctypes.c_float(B2), ctypes.c_float(nu),
                                         ptr(hist), ptr(gT), ptr(g_src), ptr(g_tgt), ptr(g_nrm), ptr(g_T0), ptr(ws), nws, stream_ptr()),
                  "e2e_icp_backward")
        return (g_src if need[0] else None), g_tgt, g_nrm, g_T0, None, None, None, None, None, None, None, None


def _library_icp_autograd(src_pc, tgt_pc, tgt_normals, initial_transform, numiters, damp, dist_thresh, grad_icp, lambda_max, B, B2, nu):
    return _ICPFunction.apply(src_pc[0], tgt_pc[0], tgt_normals[0], initial_transform, numiters, damp, dist_thresh, grad_icp, lambda_max, B, B2, nu)


def point_to_plane_ICP(src_pc, tgt_pc, tgt_normals, initial_transform, numiters=20, damp=1e-8, dist_thresh=None):
    """Gauss-Newton point-to-plane ICP.  src_pc (1,Ns,3), tgt_pc / tgt_normals (1,Nt,3), initial_transform (4,4).
    Returns (transform (4,4) aligning src to tgt, chamfer_indices of the last iteration)."""
    _check_clouds(src_pc, tgt_pc, tgt_normals)
    if initial_transform.shape != (4, 4):
        raise ValueError(f"initial_transform should have shape (4, 4). Got {tuple(initial_transform.shape)}.")
    if not _needs_grad(src_pc, tgt_pc, tgt_normals, initial_transform):
        return _library_icp(src_pc, tgt_pc, tgt_normals, initial_transform, numiters, damp, dist_thresh, 0, 0.0, 1.0, 1.0, 1.0)
    return _library_icp_autograd(src_pc, tgt_pc, tgt_normals, initial_transform, numiters, damp, dist_thresh, 0, 0.0, 1.0, 1.0, 1.0)


def point_to_plane_ICP_torch(src_pc, tgt_pc, tgt_normals, initial_transform, numiters=20, damp=1e-8, dist_thresh=None):
    """The same iteration as torch ops around the kNN kernel (autograd by torch): the reference the library's reverse sweep is
    tested against."""
    _check_clouds(src_pc, tgt_pc, tgt_normals)
    T = initial_transform
    cur = _apply(T, src_pc[0])
    idx = None
    for _ in range(numiters):
        A, b, idx = gauss_newton_solve(cur.unsqueeze(0), tgt_pc, tgt_normals, dist_thresh)
        step = se3_exp(solve_linear_system(A, b, damp)[:, 0])
        cur = _apply(step, cur)
        T = step @ T
    return T, idx


def point_to_plane_gradICP(src_pc, tgt_pc, tgt_normals, initial_transform, numiters=20, damp=1e-8, dist_thresh=None,
                           lambda_max=2.0, B=1.0, B2=1.0, nu=200.0):
    """GradICP (gradSLAM, ICRA 2020, section 3.2): Levenberg-Marquardt whose accept / reject and damping update are logistic
    gates of the look-ahead error, so the recovered transform is differentiable w.r.t. the clouds."""
    _check_clouds(src_pc, tgt_pc, tgt_normals)
    if initial_transform.shape != (4, 4):
        raise ValueError(f"initial_transform should have shape (4, 4). Got {tuple(initial_transform.shape)}.")
    if nu == 0:
        raise ValueError("nu must be non-zero")
    if not _needs_grad(src_pc, tgt_pc, tgt_normals, initial_transform):
        return _library_icp(src_pc, tgt_pc, tgt_normals, initial_transform, numiters, damp, dist_thresh, 1, lambda_max, B, B2, nu)
    return _library_icp_autograd(src_pc, tgt_pc, tgt_normals, initial_transform, numiters, damp, dist_thresh, 1, lambda_max, B, B2, nu)


def point_to_plane_gradICP_torch(src_pc, tgt_pc, tgt_normals, initial_transform, numiters=20, damp=1e-8, dist_thresh=None,
                                 lambda_max=2.0, B=1.0, B2=1.0, nu=200.0):
    """GradICP as torch ops around the kNN kernel (autograd by torch): the reference the library's reverse sweep is tested against."""
    _check_clouds(src_pc, tgt_pc, tgt_normals)
    T = initial_transform
    cur = _apply(T, src_pc[0])
    lam, lam_min = damp, damp
    idx = None
    for _ in range(numiters):
        A, b, idx = gauss_newton_solve(cur.unsqueeze(0), tgt_pc, tgt_normals, dist_thresh)
        e0 = (b * b).sum()
        xi = solve_linear_system(A, b, lam)[:, 0]
        trial = _apply(se3_exp(xi), cur)
        _, b1, _ = gauss_newton_solve(trial.unsqueeze(0), tgt_pc, tgt_normals, dist_thresh)
        e1 = (b1 * b1).sum()
        q = torch.sigmoid((e0 - e1) / nu)
        lam = lam_min + (lambda_max - lam_min) / (1.0 + B * torch.exp(-B2 * (e1 - e0) / nu))
        step = se3_exp(q * xi)
        cur = _apply(step, cur)
        T = step @ T
    return T, idx


def active_map_points(points, K, pose, H, W):
    """find_active_map_points for one batch element as a boolean mask plus pixel coordinates: map points (N,3) in front of
    the camera with pose `pose` (camera->world) whose projection falls inside the frame."""
    R, t = pose[:3, :3], pose[:3, 3]
    pc = (points - t) @ R                                     # R^T (p - t)
    z = pc[:, 2]
    u = K[0, 0] * pc[:, 0] / z + K[0, 2]
    v = K[1, 1] * pc[:, 1] / z + K[1, 2]
    inside = (z > 0) & (u > -1e-3) & (u < W - 0.999) & (v > -1e-3) & (v < H - 0.999)
    w = torch.round(u).clamp(0, W - 1).long()
    h = torch.round(v).clamp(0, H - 1).long()
    return inside, h, w
