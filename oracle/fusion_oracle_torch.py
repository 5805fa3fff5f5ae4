"""Second, independently written CPU restatement of PointFusion's map update -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED, like oracle/fusion_oracle.py (gradslam is not vendored by the reference and not installed here; see that
file's header and SURVEY.md appendix B for the semantics and the frozen decisions).  This one is written the way gradslam
itself is: torch tensors, boolean masks, `(b, n, h, w)` int64 row tensors that travel from stage to stage, and
`torch.unique(dim=0)` on `[b, h, w, 1/ccount, dist^2, n]` rows for the best-unique selection -- where the numpy oracle uses
index arrays and `np.lexsort`.  tests/test_fusion_oracles_cpu.py checks that the two restatements produce the same rows,
index maps, append order and map, bit for bit, on the synthetic sequences; tools/pin_gradslam.py replaces both as the pin
the moment `import gradslam` works.

Only tests/ may import this module.
"""
import math

import torch

F32 = torch.float32


def _sqrt(x):
    """Correctly rounded float32 square root.  torch's CPU float32 sqrt goes through MKL VML on this build and is off by one ulp
    for ~0.7 % of inputs (numpy's and CUDA's sqrtf are IEEE); the double-precision root rounded once to float32 is exact."""
    return torch.sqrt(x.double()).to(F32)


def _t(a):
    return torch.as_tensor(a, dtype=F32)


def rgbd_maps(depth, rgb, K, pose, sigma):
    """gradslam.structures.RGBDImages derived maps of one frame + PointFusion's alpha (fusionutils.get_alpha)."""
    depth, K, pose = _t(depth), _t(K), _t(pose)
    H, W = depth.shape
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    one = torch.tensor(1.0)
    ifx, ify, icx, icy = one / fx, one / fy, -(cx / fx), -(cy / fy)
    vv, uu = torch.meshgrid(torch.arange(H, dtype=F32), torch.arange(W, dtype=F32), indexing="ij")
    valid = depth > 0
    m = valid.to(F32)
    V = torch.stack([(uu * ifx + icx) * depth * m, (vv * ify + icy) * depth * m, depth * m], -1)
    dh, dv = torch.zeros_like(V), torch.zeros_like(V)
    dh[:, :-1] = V[:, 1:] - V[:, :-1]
    dv[:-1] = V[1:] - V[:-1]
    n = torch.stack([dh[..., 1] * dv[..., 2] - dh[..., 2] * dv[..., 1],
                     dh[..., 2] * dv[..., 0] - dh[..., 0] * dv[..., 2],
                     dh[..., 0] * dv[..., 1] - dh[..., 1] * dv[..., 0]], -1)
    norm = _sqrt(n[..., 0] * n[..., 0] + n[..., 1] * n[..., 1] + n[..., 2] * n[..., 2])     # left to right
    norm = torch.where(norm == 0, torch.ones_like(norm), norm)
    N = n / norm[..., None] * m[..., None]
    R, t = pose[:3, :3], pose[:3, 3]
    Vg = torch.stack([R[i, 0] * V[..., 0] + R[i, 1] * V[..., 1] + R[i, 2] * V[..., 2] + t[i] for i in range(3)], -1) * m[..., None]
    Ng = torch.stack([R[i, 0] * N[..., 0] + R[i, 1] * N[..., 1] + R[i, 2] * N[..., 2] for i in range(3)], -1)
    c = torch.tensor(2.0) * torch.tensor(float(sigma), dtype=F32) * torch.tensor(float(sigma), dtype=F32)
    arg = -((V[..., 0] * V[..., 0] + V[..., 1] * V[..., 1]) / c)
    alpha = torch.exp(arg.double()).to(F32)
    return dict(vertex=V, normal=N, vertex_g=Vg, normal_g=Ng, alpha=alpha, valid=valid)


def find_active_map_points(points, K, pose, H, W, b=0):
    """fusionutils.find_active_map_points: pc2im_bnhw rows (b, n, h, w), ordered by n."""
    points, K, pose = _t(points), _t(K), _t(pose)
    if points.shape[0] == 0:
        return torch.zeros(0, 4, dtype=torch.int64)
    Rinv = pose[:3, :3].t().contiguous()
    t = pose[:3, 3]
    tinv = -(Rinv[:, 0] * t[0] + Rinv[:, 1] * t[1] + Rinv[:, 2] * t[2])
    pc = torch.stack([Rinv[i, 0] * points[:, 0] + Rinv[i, 1] * points[:, 1] + Rinv[i, 2] * points[:, 2] + tinv[i] for i in range(3)], -1)
    hom = torch.stack([K[i, 0] * pc[:, 0] + K[i, 1] * pc[:, 1] + K[i, 2] * pc[:, 2] + K[i, 3] for i in range(3)], -1)
    u, v = hom[:, 0] / hom[:, 2], hom[:, 1] / hom[:, 2]
    lo = torch.tensor(-1e-3, dtype=F32)
    in_frame = (pc[:, 2] > 0) & (u > lo) & (u < torch.tensor(W - 0.999, dtype=F32)) & (v > lo) & (v < torch.tensor(H - 0.999, dtype=F32))
    n = torch.nonzero(in_frame)[:, 0]
    w = torch.round(u[n]).clamp(0, W - 1).long()            # torch.round: half to even
    h = torch.round(v[n]).clamp(0, H - 1).long()
    return torch.stack([torch.full_like(n, b), n, h, w], 1)


def find_similar_map_points(points, normals, maps, rows, dist_th, dot_th):
    """Keep the rows whose map point is close to, and similarly oriented as, the live vertex it projects onto."""
    points, normals = _t(points), _t(normals)
    n, h, w = rows[:, 1], rows[:, 2], rows[:, 3]
    d = maps["vertex_g"][h, w] - points[n]
    dist2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
    ng, nm = maps["normal_g"][h, w], normals[n]
    dot = ng[:, 0] * nm[:, 0] + ng[:, 1] * nm[:, 1] + ng[:, 2] * nm[:, 2]
    keep = (_sqrt(dist2) < torch.tensor(dist_th, dtype=F32)) & (dot > torch.tensor(dot_th, dtype=F32))
    return rows[keep], dist2[keep]


def find_best_unique_correspondences(ccount, rows, dist2):
    """One map point per live pixel: sort `[b, h, w, 1/(c + 1e-20), dist^2, n]` rows with torch.unique(dim=0) (lexicographic) and
    keep the first row of every (b, h, w) group.  float64 rows hold the int64 indices and the float32 keys exactly."""
    if rows.shape[0] == 0:
        return rows
    ccount = _t(ccount).reshape(-1)
    inv_c = torch.tensor(1.0) / (ccount[rows[:, 1]] + torch.tensor(1e-20, dtype=F32))
    table = torch.stack([rows[:, 0].double(), rows[:, 2].double(), rows[:, 3].double(), inv_c.double(), dist2.double(), rows[:, 1].double()], 1)
    table = torch.unique(table, dim=0)                                   # sorted lexicographically, duplicates impossible (n is a column)
    first = torch.ones(table.shape[0], dtype=torch.bool)
    first[1:] = (table[1:, :3] != table[:-1, :3]).any(1)
    best = table[first]
    out = torch.stack([best[:, 0], best[:, 5], best[:, 1], best[:, 2]], 1).long()
    return out[torch.argsort(out[:, 1], stable=True)]


def fuse_with_map(points, normals, colors, ccount, rgb, maps, rows):
    points, normals, colors, ccount, rgb = _t(points).clone(), _t(normals).clone(), _t(colors).clone(), _t(ccount).reshape(-1).clone(), _t(rgb)
    H, W = maps["alpha"].shape
    n, h, w = rows[:, 1], rows[:, 2], rows[:, 3]
    c, a = ccount[n][:, None], maps["alpha"][h, w][:, None]
    den = c + a
    points[n] = (c * points[n] + a * maps["vertex_g"][h, w]) / den
    normals[n] = (c * normals[n] + a * maps["normal_g"][h, w]) / den
    colors[n] = (c * colors[n] + a * rgb[h, w]) / den
    ccount[n] = den[:, 0]
    matched = torch.zeros(H, W, dtype=torch.bool)
    matched[h, w] = True
    new = maps["valid"] & ~matched
    return (torch.cat([points, maps["vertex_g"][new]]), torch.cat([normals, maps["normal_g"][new]]), torch.cat([colors, rgb[new]]),
            torch.cat([ccount, maps["alpha"][new]]), new)


class PointFusionOracleTorch:
    def __init__(self, dist_th=0.05, angle_th=20, sigma=0.6):
        self.dist_th, self.sigma = dist_th, sigma
        self.dot_th = math.cos(angle_th * math.pi / 180.0)
        z = torch.zeros(0, 3, dtype=F32)
        self.points, self.normals, self.colors, self.ccount = z, z.clone(), z.clone(), torch.zeros(0, dtype=F32)

    def step(self, depth, rgb, K, pose):
        maps = rgbd_maps(depth, rgb, K, pose, self.sigma)
        H, W = maps["alpha"].shape
        active = find_active_map_points(self.points, K, pose, H, W)
        similar, dist2 = find_similar_map_points(self.points, self.normals, maps, active, self.dist_th, self.dot_th)
        rows = find_best_unique_correspondences(self.ccount, similar, dist2)
        self.points, self.normals, self.colors, self.ccount, new = fuse_with_map(self.points, self.normals, self.colors, self.ccount,
                                                                                 rgb, maps, rows)
        return dict(maps=maps, active=active, rows=rows, appended=new)
