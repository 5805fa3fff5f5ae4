"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float64 arithmetic on float32 inputs) of the point-to-plane ICP /
GradICP odometry that gradslam's PointFusion runs when odom != "gt" (call sites
train_depth.py:111-116, 378-381, online_adaption.py:362-363; configs/config.yaml:30-34).

PARITY UNPINNED: gradslam (gradslam/odometry/icputils.py, gradslam/geometry/se3utils.py, v0.1.0 layout) is a
third-party, un-vendored, un-pinned dependency of the reference and is not installed here; the reference holds no test
or fixture for it.  This file restates the PUBLISHED algorithm (gradSLAM, Jatavallabhula et al., ICRA 2020, section
3.2 "differentiable optimisation": Gauss-Newton point-to-plane ICP; GradICP = Levenberg-Marquardt whose discrete
accept / reject and damping update are replaced by logistic gates) with these conventions FROZEN:

  * correspondences: K = 1 nearest target point of every (already transformed) source point, squared L2 (knn1 of
    oracle/fusion_oracle.py); `dist_thresh`, if given, keeps pairs with squared distance < dist_thresh;
  * residual b_i = n_i . (d_i - s_i);  Jacobian row A_i = [n_i, s_i x n_i]  (translation first, rotation second);
  * step xi = (A^T A + damp * I)^-1 A^T b ; xi = [v, omega] ; T_step = se3_exp(xi) ; source <- T_step source ;
    transform <- T_step @ transform ; numiters iterations, no early exit;
  * se3_exp: R = I + sin(t)/t K + (1 - cos t)/t^2 K^2, V = I + (1 - cos t)/t^2 K + (t - sin t)/t^3 K^2, translation V v,
    with R = I + K, V = I + K / 2 when t^2 < 1e-12;
  * GradICP: e0 = |b|^2 at the current source, e1 = |b|^2 after the trial step; gate q = 1 / (1 + exp(-(e0 - e1) / nu)),
    applied step = se3_exp(q * xi); next damping = lambda_min + (lambda_max - lambda_min) / (1 + B * exp(-B2 * (e1 - e0) / nu))
    with lambda_min = the `damp` argument (also the first iteration's damping).
"""
import numpy as np

from oracle.fusion_oracle import knn1


def se3_exp(xi):
    xi = np.asarray(xi, np.float64)
    v, w = xi[:3], xi[3:]
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], np.float64)
    t2 = float(w @ w)
    if t2 < 1e-12:
        R, V = np.eye(3) + K, np.eye(3) + 0.5 * K
    else:
        t = np.sqrt(t2)
        R = np.eye(3) + np.sin(t) / t * K + (1 - np.cos(t)) / t2 * (K @ K)
        V = np.eye(3) + (1 - np.cos(t)) / t2 * K + (t - np.sin(t)) / (t2 * t) * (K @ K)
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, V @ v
    return T


def linearize(src, tgt, tgt_normals, dist_thresh=None):
    d2, idx = knn1(src.astype(np.float32), tgt.astype(np.float32))
    keep = np.ones(len(src), bool) if dist_thresh is None else d2 < np.float32(dist_thresh)
    s = src[keep].astype(np.float64)
    d, n = tgt[idx[keep]].astype(np.float64), tgt_normals[idx[keep]].astype(np.float64)
    A = np.concatenate([n, np.cross(s, n)], 1)
    b = (n * (d - s)).sum(1)
    return A, b, idx


def solve(A, b, damp):
    return np.linalg.solve(A.T @ A + damp * np.eye(6), A.T @ b)


def _apply(T, pts):
    return (pts.astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)


def point_to_plane_icp(src, tgt, tgt_normals, initial_transform, numiters=20, damp=1e-8, dist_thresh=None):
    T = np.asarray(initial_transform, np.float64).copy()
    cur = _apply(T, src)
    errs = []
    for _ in range(numiters):
        A, b, idx = linearize(cur, tgt, tgt_normals, dist_thresh)
        errs.append(float(b @ b))
        step = se3_exp(solve(A, b, damp))
        cur = _apply(step, cur)
        T = step @ T
    return T, idx, errs


def point_to_plane_gradicp(src, tgt, tgt_normals, initial_transform, numiters=20, damp=1e-8, dist_thresh=None,
                           lambda_max=2.0, B=1.0, B2=1.0, nu=200.0):
    T = np.asarray(initial_transform, np.float64).copy()
    cur = _apply(T, src)
    lam, lam_min = float(damp), float(damp)
    errs = []
    for _ in range(numiters):
        A, b, idx = linearize(cur, tgt, tgt_normals, dist_thresh)
        e0 = float(b @ b)
        xi = solve(A, b, lam)
        trial = _apply(se3_exp(xi), cur)
        _, b1, _ = linearize(trial, tgt, tgt_normals, dist_thresh)
        e1 = float(b1 @ b1)
        q = 1.0 / (1.0 + np.exp(-(e0 - e1) / nu))
        lam = lam_min + (lambda_max - lam_min) / (1.0 + B * np.exp(-B2 * (e1 - e0) / nu))
        step = se3_exp(q * xi)
        cur = _apply(step, cur)
        T = step @ T
        errs.append(e0)
    return T, idx, errs
