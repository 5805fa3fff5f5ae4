"""CPU oracle for PointFusion (gradslam semantics) and K=1 nearest neighbour -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The arithmetic restated here lives in third-party packages that the reference imports
but does not vendor or pin, and that are not installed in the build container:
  * gradslam (README.md:5-24 offers PyPI / git HEAD / local clone; the import
    `gradslam.slam.fusionutils.find_active_map_points` at online_adaption.py:35 matches the v0.1.0 layout):
    structures/rgbdimages.py (vertex / normal maps), slam/fusionutils.py (find_active_map_points,
    find_similar_map_points, find_best_unique_correspondences, fuse_with_map, get_alpha),
    geometry/geometryutils.py (transform_pointcloud), structures/pointclouds.py (append_points);
  * chamferdist (loss/losses.py:42 cites commit 255b7108...): chamfer.knn_points, K = 1.
No test, fixture or golden vector in the reference exercises either, so nothing pins this restatement to
their actual output.  It follows SURVEY.md appendix B line by line; every decision the survey lists as
"to freeze" is frozen below and marked [FROZEN].  The reference's own call sites anchor the interface:
slam/custom_slam.py:26-34, online_adaption.py:329-366, 457-471, 638-645, train_depth.py:111-118, 263-267.

Written in plain numpy, vectorised with masks / lexsort (the way gradslam does it with torch.unique), and
deliberately NOT shaped like the CUDA design (per-pixel atomicMin keys), so that agreement of the integer
outputs (index map, correspondence rows, append order) is meaningful.  All floating point is float32 with
one rounding per written operation (numpy never fuses), in the written order; the CUDA kernels mirror
that order with __fmul_rn/__fadd_rn, which is what makes bit-exact integer outputs attainable.
"""
import math

import numpy as np

f32 = np.float32


def _f(a):
    return np.asarray(a, dtype=np.float32)


def inverse_pose(pose):
    """[FROZEN] rigid inverse: R^T, -R^T t, each dot product accumulated left to right."""
    pose = _f(pose)
    R, t = pose[:3, :3], pose[:3, 3]
    Rinv = R.T.copy()
    tinv = -((Rinv[:, 0] * t[0] + Rinv[:, 1] * t[1]) + Rinv[:, 2] * t[2])
    return Rinv, tinv.astype(np.float32)


def rgbd_maps(depth, rgb, K, pose, sigma):
    """RGBDImages-derived maps of one frame + PointFusion's per-pixel confidence.

    depth (H,W), rgb (H,W,3), K (4,4), pose (4,4) camera->world.
    Returns dict(vertex (H,W,3) local, normal local, vertex_g, normal_g, alpha (H,W), valid (H,W) bool)."""
    depth, K, pose = _f(depth), _f(K), _f(pose)
    H, W = depth.shape
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    # [FROZEN] closed-form inverse intrinsics (gradslam projutils.inverse_intrinsics)
    ifx, ify = f32(1.0) / fx, f32(1.0) / fy
    icx, icy = -(cx / fx), -(cy / fy)
    u = np.arange(W, dtype=np.float32)[None, :].repeat(H, 0)
    v = np.arange(H, dtype=np.float32)[:, None].repeat(W, 1)
    valid = depth > 0                                           # valid_depth_mask
    m = valid.astype(np.float32)
    X = ((u * ifx) + icx) * depth                               # [FROZEN] (u*ifx + icx) * d
    Y = ((v * ify) + icy) * depth
    V = np.stack([X * m, Y * m, depth * m], -1).astype(np.float32)   # masked vertex map
    dh = np.zeros_like(V)
    dv = np.zeros_like(V)
    dh[:, :-1] = V[:, 1:] - V[:, :-1]                           # forward differences; last col / row stay 0
    dv[:-1, :] = V[1:, :] - V[:-1, :]
    n = np.stack([dh[..., 1] * dv[..., 2] - dh[..., 2] * dv[..., 1],
                  dh[..., 2] * dv[..., 0] - dh[..., 0] * dv[..., 2],
                  dh[..., 0] * dv[..., 1] - dh[..., 1] * dv[..., 0]], -1).astype(np.float32)
    norm = np.sqrt((n[..., 0] * n[..., 0] + n[..., 1] * n[..., 1]) + n[..., 2] * n[..., 2])
    norm = np.where(norm == 0, f32(1.0), norm).astype(np.float32)
    N = (n / norm[..., None]) * m[..., None]                    # [FROZEN] zero normal at invalid pixels
    R, t = pose[:3, :3], pose[:3, 3]
    Vg = np.stack([((R[i, 0] * V[..., 0] + R[i, 1] * V[..., 1]) + R[i, 2] * V[..., 2]) + t[i] for i in range(3)], -1)
    Vg = (Vg * m[..., None]).astype(np.float32)                 # [FROZEN] global vertex map masked again
    Ng = np.stack([(R[i, 0] * N[..., 0] + R[i, 1] * N[..., 1]) + R[i, 2] * N[..., 2] for i in range(3)], -1).astype(np.float32)
    c = f32(2.0) * f32(sigma) * f32(sigma)
    arg = -(((V[..., 0] * V[..., 0]) + (V[..., 1] * V[..., 1])) / c)                # [FROZEN] LOCAL x, y
    alpha = np.exp(arg.astype(np.float64)).astype(np.float32)   # [FROZEN] float32 argument, exp in double, one rounding
    return dict(vertex=V, normal=N.astype(np.float32), vertex_g=Vg, normal_g=Ng, alpha=alpha, valid=valid)


def find_active_map_points(points, K, pose, H, W, b=0):
    """gradslam.slam.fusionutils.find_active_map_points (imported by the reference at online_adaption.py:35; SURVEY.md appendix B
    step 1): rows (b, n, h, w) int64 of the map points in front of the live camera that project into the frame, ordered by n."""
    points, K = _f(points), _f(K)
    if points.shape[0] == 0:
        return np.zeros((0, 4), np.int64)
    Rinv, tinv = inverse_pose(pose)
    pc = np.stack([((Rinv[i, 0] * points[:, 0] + Rinv[i, 1] * points[:, 1]) + Rinv[i, 2] * points[:, 2]) + tinv[i]
                   for i in range(3)], -1).astype(np.float32)
    front = pc[:, 2] > 0
    hom = np.stack([((K[i, 0] * pc[:, 0] + K[i, 1] * pc[:, 1]) + K[i, 2] * pc[:, 2]) + K[i, 3] for i in range(3)], -1)
    with np.errstate(divide="ignore", invalid="ignore"):
        uu = (hom[:, 0] / hom[:, 2]).astype(np.float32)
        vv = (hom[:, 1] / hom[:, 2]).astype(np.float32)
    in_frame = (uu > f32(-1e-3)) & (uu < f32(W - 0.999)) & (vv > f32(-1e-3)) & (vv < f32(H - 0.999)) & front   # [FROZEN]
    n_idx = np.nonzero(in_frame)[0]
    w = np.clip(np.rint(uu[n_idx]), 0, W - 1).astype(np.int64)  # rint = round half to even, like torch.round [FROZEN]
    h = np.clip(np.rint(vv[n_idx]), 0, H - 1).astype(np.int64)
    return np.stack([np.full_like(n_idx, b), n_idx, h, w], -1).astype(np.int64)


def find_correspondences(points, normals, ccount, K, pose, maps, dist_th, dot_th):
    """Steps 1-3 of update_map_fusion.  Returns rows (M,4) int64 = (b=0, n, h, w), sorted by n -- one row
    per matched live pixel -- and the same information as index_map (H,W) int64 (-1 = no match)."""
    points, normals, ccount, K = _f(points), _f(normals), _f(ccount).reshape(-1), _f(K)
    Vg, Ng = maps["vertex_g"], maps["normal_g"]
    H, W = Vg.shape[:2]
    index_map = np.full((H, W), -1, dtype=np.int64)
    if points.shape[0] == 0:
        return np.zeros((0, 4), np.int64), index_map
    # -- 1. active map points (find_active_map_points) --------------------------------------------
    active = find_active_map_points(points, K, pose, H, W)
    n_idx, h, w = active[:, 1], active[:, 2], active[:, 3]
    # -- 2. similar points (find_similar_map_points) ----------------------------------------------
    d = Vg[h, w] - points[n_idx]
    dist2 = ((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]).astype(np.float32)
    dist = np.sqrt(dist2)
    nm = normals[n_idx]
    dot = ((Ng[h, w, 0] * nm[:, 0] + Ng[h, w, 1] * nm[:, 1]) + Ng[h, w, 2] * nm[:, 2]).astype(np.float32)
    keep = (dist < f32(dist_th)) & (dot > f32(dot_th))          # strict inequalities [FROZEN]
    n_idx, h, w, dist2 = n_idx[keep], h[keep], w[keep], dist2[keep]
    # -- 3. best unique correspondence per pixel (find_best_unique_correspondences) -----------------
    inv_c = (f32(1.0) / (ccount[n_idx] + f32(1e-20))).astype(np.float32)
    order = np.lexsort((n_idx, dist2, inv_c, w, h))             # sort rows (h, w, 1/c, d^2, n) lexicographically
    n_idx, h, w = n_idx[order], h[order], w[order]
    first = np.ones(len(order), bool)
    first[1:] = (h[1:] != h[:-1]) | (w[1:] != w[:-1])           # first row of every (h, w) group
    n_idx, h, w = n_idx[first], h[first], w[first]
    index_map[h, w] = n_idx
    rows = np.stack([np.zeros_like(n_idx), n_idx, h, w], -1)
    return rows[np.argsort(rows[:, 1], kind="stable")], index_map


def fuse_with_map(points, normals, colors, ccount, rgb, maps, index_map):
    """fuse_with_map: confidence-weighted merge of matched map points + append of unmatched valid pixels in
    row-major order.  Returns new (points, normals, colors, ccount) and append_slot (H,W) int64."""
    points, normals, colors, ccount, rgb = _f(points).copy(), _f(normals).copy(), _f(colors).copy(), _f(ccount).reshape(-1).copy(), _f(rgb)
    Vg, Ng, alpha, valid = maps["vertex_g"], maps["normal_g"], maps["alpha"], maps["valid"]
    h, w = np.nonzero(index_map >= 0)
    n = index_map[h, w]
    c, a = ccount[n][:, None], alpha[h, w][:, None]
    den = c + a
    points[n] = ((c * points[n]) + (a * Vg[h, w])) / den        # [FROZEN] (c*old + a*new) / (c + a)
    normals[n] = ((c * normals[n]) + (a * Ng[h, w])) / den      # [FROZEN] not re-normalised
    colors[n] = ((c * colors[n]) + (a * rgb[h, w])) / den
    ccount[n] = den[:, 0]
    new = valid & (index_map < 0)
    append_slot = np.full(index_map.shape, -1, np.int64)
    append_slot[new] = len(points) + np.arange(int(new.sum()))
    points = np.concatenate([points, Vg[new]], 0)
    normals = np.concatenate([normals, Ng[new]], 0)
    colors = np.concatenate([colors, rgb[new]], 0)
    ccount = np.concatenate([ccount, alpha[new]], 0)
    return points, normals, colors, ccount, append_slot


class PointFusionOracle:
    """PointFusion(odom='gt') restated: step() = localise with the frame's own pose, then update_map_fusion."""

    def __init__(self, dist_th=0.05, angle_th=20, sigma=0.6):
        self.dist_th, self.sigma = dist_th, sigma
        self.dot_th = math.cos(angle_th * math.pi / 180.0)
        self.points = np.zeros((0, 3), np.float32)
        self.normals = np.zeros((0, 3), np.float32)
        self.colors = np.zeros((0, 3), np.float32)
        self.ccount = np.zeros((0,), np.float32)
        self.last = None

    def step(self, depth, rgb, K, pose):
        maps = rgbd_maps(depth, rgb, K, pose, self.sigma)
        rows, index_map = find_correspondences(self.points, self.normals, self.ccount, K, pose, maps, self.dist_th, self.dot_th)
        self.points, self.normals, self.colors, self.ccount, slot = fuse_with_map(
            self.points, self.normals, self.colors, self.ccount, rgb, maps, index_map)
        self.last = dict(maps=maps, rows=rows, index_map=index_map, append_slot=slot)
        return self.last


def transform_pointcloud(points, T):
    """gradslam.geometry.geometryutils.transform_pointcloud: R p + t   [FROZEN] left-to-right accumulation."""
    points, T = _f(points), _f(T)
    return np.stack([((T[i, 0] * points[:, 0] + T[i, 1] * points[:, 1]) + T[i, 2] * points[:, 2]) + T[i, 3]
                     for i in range(3)], -1).astype(np.float32)


def knn1(query, ref, chunk=2048):
    """chamferdist.chamfer.knn_points with K = 1: squared L2 distance to, and index of, the nearest reference
    point for every query point.  [FROZEN] d^2 = ((dx*dx + dy*dy) + dz*dz) in float32, first minimum wins."""
    query, ref = _f(query), _f(ref)
    P1 = query.shape[0]
    dist2 = np.empty(P1, np.float32)
    idx = np.empty(P1, np.int64)
    for s in range(0, P1, chunk):
        q = query[s:s + chunk, None, :]
        d = q - ref[None, :, :]
        d2 = ((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]).astype(np.float32)
        i = np.argmin(d2, axis=1)                               # first occurrence of the minimum
        idx[s:s + chunk] = i
        dist2[s:s + chunk] = d2[np.arange(d2.shape[0]), i]
    return dist2, idx


def synthetic_room_sequence(L, H, W, seed=0):
    """Config C3: an analytic box room ray-cast to L depth + colour frames along a smooth trajectory
    (1-3 cm and <= 1 degree per frame), with GT poses.  Returns depth (L,H,W), rgb (L,H,W,3), K, poses."""
    rng = np.random.default_rng(seed)
    fx = fy = 525.0 * W / 640.0
    cx, cy = W / 2 - 0.5, H / 2 - 0.5
    K = np.eye(4, dtype=np.float32)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2] = fx, fy, cx, cy
    half = np.array([2.0, 1.4, 3.5])                            # room half extents (m), camera starts near the centre
    uu, vv = np.meshgrid(np.arange(W), np.arange(H))
    rays = np.stack([(uu - cx) / fx, (vv - cy) / fy, np.ones_like(uu, dtype=np.float64)], -1)
    depth = np.empty((L, H, W), np.float32)
    rgb = np.empty((L, H, W, 3), np.float32)
    poses = np.empty((L, 4, 4), np.float32)
    pos = np.array([0.1, -0.1, 0.0])
    yaw = pitch = 0.0
    for s in range(L):
        cyaw, syaw, cp, sp = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch)
        Ry = np.array([[cyaw, 0, syaw], [0, 1, 0], [-syaw, 0, cyaw]])
        Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
        R = Ry @ Rx
        dirs = rays @ R.T
        with np.errstate(divide="ignore"):
            tpos = (half - pos) / dirs
            tneg = (-half - pos) / dirs
        tt = np.where(dirs > 0, tpos, tneg)
        tt = np.where(np.isfinite(tt) & (tt > 0), tt, np.inf)
        thit = tt.min(-1)
        wall = tt.argmin(-1)
        hit = pos + dirs * thit[..., None]
        depth[s] = thit.astype(np.float32)                      # z-depth: rays have unit z in the camera frame
        tex = 0.5 + 0.25 * np.sin(3.0 * hit[..., 0] + wall) + 0.25 * np.cos(2.5 * hit[..., 1] + 1.3 * hit[..., 2])
        rgb[s] = np.clip(np.stack([tex, 0.8 * tex + 0.1 * np.sin(hit[..., 2]), 1.0 - tex], -1), 0, 1)
        poses[s] = np.eye(4)
        poses[s, :3, :3] = R
        poses[s, :3, 3] = pos
        pos = pos + np.array([0.02, 0.003, 0.015]) + rng.normal(0, 0.002, 3)
        yaw += math.radians(0.6)
        pitch += math.radians(0.1) * math.sin(s / 5.0)
    return depth, rgb, K, poses
