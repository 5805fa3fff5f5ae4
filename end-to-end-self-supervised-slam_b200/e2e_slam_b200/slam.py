"""gradslam-compatible surface for the part of gradslam the reference drives:
`RGBDImages`, `Pointclouds`, `PointFusion` (.step / __call__), `transform_pointcloud`, plus the reference's own
`image_recover_slam` (slam/custom_slam.py:6-35).  Only the attributes the reference touches are provided:
    RGBDImages(rgb_image, depth_image, intrinsics, poses)[:, s] / .to / .detach / .shape / .poses (settable) /
        .rgb_image / .depth_image / .intrinsics
    Pointclouds(device=).points_list / .colors_list / .normals_list / .features_list / .has_points / .detach() /
        [b] / len()
    PointFusion(odom=, dist_th=, angle_th=, sigma=, numiters=, device=).step(pointclouds, live_frame,
        prev_frame=None, inplace=False) -> (pointclouds, poses);  slam(frames) -> (pointclouds, poses)
Fusion semantics are those of oracle/fusion_oracle.py (gradslam as restated in SURVEY.md appendix B).
The map size lives on the device; a step launches 7 kernels and never synchronises with the host.
Reading `points_list` does (it has to know N).

Odometry: PointFusion.step localises with the frame's own pose when `odom == "gt"` or `prev_frame is None`
-- exactly gradslam's rule.  Otherwise (`odom` "icp" / "gradicp") the live frame, thinned
to every dsratio-th pixel and placed with the previous pose, is aligned to the active map points seen from the previous
frame (odometry.point_to_plane_ICP / point_to_plane_gradICP) and the result is composed with the previous pose.
"""
import ctypes
import math

import torch

from ._lib import check, f32, lib, ptr, stream_ptr


class RGBDImages:
    """Channels-last RGB-D sequence container: rgb (B,L,H,W,3), depth (B,L,H,W,1), intrinsics (B,1,4,4),
    poses (B,L,4,4) camera->world (may be None)."""

    def __init__(self, rgb_image, depth_image, intrinsics, poses=None, channels_first=False, device=None):
        for name, t, nd in (("rgb_image", rgb_image, 5), ("depth_image", depth_image, 5), ("intrinsics", intrinsics, 4)):
            if not torch.is_tensor(t):
                raise TypeError(f"Expected {name} to be of type tensor. Got {type(t)}.")
            if t.dim() != nd:
                raise ValueError(f"{name} should have {nd} dimensions. Got {t.dim()}.")
        if channels_first:
            rgb_image, depth_image = rgb_image.permute(0, 1, 3, 4, 2), depth_image.permute(0, 1, 3, 4, 2)
        if rgb_image.shape[:4] != depth_image.shape[:4] or rgb_image.shape[-1] != 3 or depth_image.shape[-1] != 1:
            raise ValueError(f"rgb_image {tuple(rgb_image.shape)} and depth_image {tuple(depth_image.shape)} are inconsistent")
        if intrinsics.shape != (rgb_image.shape[0], 1, 4, 4):
            raise ValueError(f"intrinsics should be ({rgb_image.shape[0]},1,4,4). Got {tuple(intrinsics.shape)}.")
        if poses is not None and (not torch.is_tensor(poses) or poses.shape != (*rgb_image.shape[:2], 4, 4)):
            raise ValueError(f"poses should be {(*rgb_image.shape[:2], 4, 4)}")
        self._rgb, self._depth, self._K, self._poses = rgb_image, depth_image, intrinsics, poses
        if device is not None:
            self.to(device)

    rgb_image = property(lambda self: self._rgb)
    depth_image = property(lambda self: self._depth)
    intrinsics = property(lambda self: self._K)
    device = property(lambda self: self._rgb.device)

    @property
    def poses(self):
        return self._poses

    @poses.setter
    def poses(self, value):
        if value is not None and (not torch.is_tensor(value) or value.shape != (*self._rgb.shape[:2], 4, 4)):
            raise ValueError(f"poses should be {(*self._rgb.shape[:2], 4, 4)}")
        self._poses = value

    @property
    def shape(self):
        return tuple(self._rgb.shape[:4])

    def __len__(self):
        return self._rgb.shape[0]

    def __getitem__(self, index):
        if not isinstance(index, tuple):
            index = (index,)
        if len(index) > 2:
            raise IndexError("RGBDImages supports indexing of the batch and sequence dimensions only")
        dims = self._rgb.shape[:2]

        def _keep(i, size):                                   # ints keep their dimension; negative ints count from the end
            if not isinstance(i, int):
                return i
            j = i + size if i < 0 else i
            if not (0 <= j < size):
                raise IndexError(f"index {i} is out of bounds for dimension with size {size}")
            return slice(j, j + 1)
        keep = tuple(_keep(i, dims[d]) for d, i in enumerate(index))
        bidx = keep[0]
        return RGBDImages(self._rgb[keep], self._depth[keep], self._K[bidx],
                          None if self._poses is None else self._poses[keep])

    def to(self, device):
        self._rgb, self._depth, self._K = self._rgb.to(device), self._depth.to(device), self._K.to(device)
        if self._poses is not None:
            self._poses = self._poses.to(device)
        return self

    def detach(self):
        return RGBDImages(self._rgb.detach(), self._depth.detach(), self._K.detach(),
                          None if self._poses is None else self._poses.detach())

    def clone(self):
        return RGBDImages(self._rgb.clone(), self._depth.clone(), self._K.clone(),
                          None if self._poses is None else self._poses.clone())

    def plotly(self, *a, **k):
        raise NotImplementedError("visualisation is out of scope (plotly is not a dependency of this package)")


class _Map:
    """One batch element's map: structure of arrays with spare capacity; N lives on the device."""

    __slots__ = ("pts", "nrm", "col", "cc", "n_dev", "n_upper", "n_host")

    def __init__(self, device):
        z = dict(dtype=torch.float32, device=device)
        self.pts, self.nrm, self.col = torch.empty(0, 3, **z), torch.empty(0, 3, **z), torch.empty(0, 3, **z)
        self.cc = torch.empty(0, **z)
        self.n_dev = torch.zeros(1, dtype=torch.int64, device=device)
        self.n_upper = 0       # host-side upper bound of N (sizes grids; tightened when N is read)
        self.n_host = 0        # last N seen by the host (exact only right after count())

    def count(self):
        self.n_host = int(self.n_dev.item())       # the one place the host synchronises
        self.n_upper = self.n_host
        return self.n_host

    def copy(self, detach=False):
        m = _Map(self.pts.device)
        f = (lambda t: t.detach()) if detach else (lambda t: t)
        m.pts, m.nrm, m.col, m.cc = f(self.pts), f(self.nrm), f(self.col), f(self.cc)
        m.n_dev, m.n_upper, m.n_host = self.n_dev, self.n_upper, self.n_host
        return m


class Pointclouds:
    """Batch of point clouds with per-point normals, colours and a scalar feature (PointFusion's confidence
    count).  Lists may be given at construction (gradslam-style) or the object starts empty."""

    def __init__(self, points=None, normals=None, colors=None, features=None, device=None):
        self.device = torch.device(device) if device is not None else (points[0].device if points else torch.device("cuda"))
        self._maps = []
        if points is not None:
            for b, p in enumerate(points):
                m = _Map(self.device)
                n = p.shape[0]
                m.pts = f32(p.to(self.device), "points").contiguous()
                m.nrm = normals[b].to(self.device).contiguous() if normals is not None else torch.zeros_like(m.pts)
                m.col = colors[b].to(self.device).contiguous() if colors is not None else torch.zeros_like(m.pts)
                m.cc = (features[b].to(self.device).reshape(-1).contiguous() if features is not None
                        else torch.ones(n, dtype=torch.float32, device=self.device))
                m.n_dev = torch.full((1,), n, dtype=torch.int64, device=self.device)
                m.n_upper = m.n_host = n
                self._maps.append(m)

    def __len__(self):
        return len(self._maps)

    @property
    def has_points(self):
        return any(m.n_upper > 0 for m in self._maps)

    def _list(self, attr):
        out = []
        for m in self._maps:
            n = m.count()
            out.append(getattr(m, attr)[:n])
        return out

    points_list = property(lambda self: self._list("pts"))
    normals_list = property(lambda self: self._list("nrm"))
    colors_list = property(lambda self: self._list("col"))
    features_list = property(lambda self: [c.unsqueeze(-1) for c in self._list("cc")])

    @property
    def num_points_per_pointcloud(self):
        return torch.tensor([m.count() for m in self._maps], dtype=torch.int64)

    def __getitem__(self, b):
        pc = Pointclouds(device=self.device)
        pc._maps = [self._maps[b]] if isinstance(b, int) else self._maps[b]
        return pc

    def detach(self):
        pc = Pointclouds(device=self.device)
        pc._maps = [m.copy(detach=True) for m in self._maps]
        return pc

    def clone(self):
        pc = Pointclouds(device=self.device)
        for m in self._maps:
            c = m.copy()
            c.pts, c.nrm, c.col, c.cc, c.n_dev = m.pts.clone(), m.nrm.clone(), m.col.clone(), m.cc.clone(), m.n_dev.clone()
            pc._maps.append(c)
        return pc

    def to(self, device):
        if torch.device(device) != self.device:
            raise NotImplementedError("Pointclouds live on the device they were created on")
        return self

    def plotly(self, *a, **k):
        raise NotImplementedError("visualisation is out of scope (plotly is not a dependency of this package)")


def _frame_maps(depth, K, pose, sigma):
    """vertex_g, normal_g [H,W,3], alpha [H,W], valid [H,W] uint8 of one frame (e2e_rgbd_maps)."""
    H, W = depth.shape
    dev = depth.device
    vg = torch.empty(H, W, 3, dtype=torch.float32, device=dev)
    ng = torch.empty(H, W, 3, dtype=torch.float32, device=dev)
    alpha = torch.empty(H, W, dtype=torch.float32, device=dev)
    valid = torch.empty(H, W, dtype=torch.uint8, device=dev)
    check(lib().e2e_rgbd_maps(ptr(depth), None, ptr(K), ptr(pose), H, W, ctypes.c_float(sigma), ptr(vg), ptr(ng), ptr(alpha),
                              ptr(valid), stream_ptr()), "e2e_rgbd_maps")
    return vg, ng, alpha, valid


class _FusionStep(torch.autograd.Function):
    """One update_map_fusion for one batch element.  Returns new (points, normals, colors, ccount, n_dev,
    index_map, append_slot); the buffers have spare capacity, N is n_dev."""

    @staticmethod
    def forward(ctx, depth, rgb, K, pose, old_pts, old_nrm, old_col, old_cc, n_dev, n_upper, dist_th, dot_th, sigma, in_place):
        H, W = depth.shape
        dev = depth.device
        with torch.cuda.device(dev):
            vg, ng, alpha, valid = _frame_maps(depth, K, pose, sigma)
            keys = torch.empty(H, W, dtype=torch.int64, device=dev)
            cand = torch.empty(max(n_upper, 1), dtype=torch.int32, device=dev)
            index_map = torch.empty(H, W, dtype=torch.int64, device=dev)
            check(lib().e2e_fusion_associate(ptr(old_pts), ptr(old_nrm), ptr(old_cc), ptr(n_dev), n_upper, ptr(K), ptr(pose),
                                             ptr(vg), ptr(ng), H, W, ctypes.c_float(dist_th), ctypes.c_float(dot_th),
                                             ptr(keys), ptr(cand), ptr(index_map), stream_ptr()), "e2e_fusion_associate")
            need = n_upper + H * W
            cap_old = old_pts.shape[0]
            if in_place and cap_old >= need:
                pts, nrm, col, cc = old_pts, old_nrm, old_col, old_cc
            else:
                cap = max(need, int(1.5 * cap_old)) if in_place else need
                z = dict(dtype=torch.float32, device=dev)
                pts, nrm, col, cc = torch.empty(cap, 3, **z), torch.empty(cap, 3, **z), torch.empty(cap, 3, **z), torch.empty(cap, **z)
                k = min(n_upper, cap_old)
                pts[:k], nrm[:k], col[:k], cc[:k] = old_pts[:k], old_nrm[:k], old_col[:k], old_cc[:k]
            append_slot = torch.empty(H, W, dtype=torch.int64, device=dev)
            n_out = torch.empty(1, dtype=torch.int64, device=dev)
            nws = lib().e2e_fusion_workspace_bytes(H, W)
            ws = torch.empty(nws, dtype=torch.uint8, device=dev)
            check(lib().e2e_fusion_merge_append(ptr(pts), ptr(nrm), ptr(col), ptr(cc), ptr(n_dev), pts.shape[0], ptr(vg), ptr(ng),
                                                ptr(rgb), ptr(alpha), ptr(valid), ptr(index_map), H, W, ptr(append_slot),
                                                ptr(n_out), ptr(ws), nws, stream_ptr()), "e2e_fusion_merge_append")
        ctx.save_for_backward(depth, K, pose, old_pts, old_col, old_cc, vg, rgb, alpha, index_map, append_slot)
        ctx.sigma = sigma
        ctx.mark_non_differentiable(nrm, n_out, index_map, append_slot)
        if in_place and pts is old_pts:
            ctx.mark_dirty(old_pts, old_nrm, old_col, old_cc)
        return pts, nrm, col, cc, n_out, index_map, append_slot

    @staticmethod
    def backward(ctx, g_pts, g_nrm, g_col, g_cc, *_):
        depth, K, pose, old_pts, old_col, old_cc, vg, rgb, alpha, index_map, append_slot = ctx.saved_tensors
        H, W = depth.shape
        dev = depth.device
        z = dict(dtype=torch.float32, device=dev)
        g_vg, g_rgb, g_alpha = torch.empty(H, W, 3, **z), torch.empty(H, W, 3, **z), torch.empty(H, W, **z)
        cap_old = old_pts.shape[0]
        need_old = any(ctx.needs_input_grad[i] for i in (4, 6, 7)) and cap_old > 0

        def passthrough(g, shape):
            if not need_old:
                return None
            return g[:cap_old].clone() if g is not None else torch.zeros(shape, **z)

        go_pts, go_col, go_cc = passthrough(g_pts, (cap_old, 3)), passthrough(g_col, (cap_old, 3)), passthrough(g_cc, (cap_old,))
        c = lambda t: None if t is None else t.contiguous()
        with torch.cuda.device(dev):
            check(lib().e2e_fusion_merge_append_bwd(ptr(c(g_pts)), ptr(c(g_col)), ptr(c(g_cc)), ptr(old_pts), ptr(old_col), ptr(old_cc),
                                                    ptr(vg), ptr(rgb), ptr(alpha), ptr(index_map), ptr(append_slot), H, W,
                                                    ptr(g_vg), ptr(g_rgb), ptr(g_alpha), ptr(go_pts), ptr(go_col), ptr(go_cc),
                                                    stream_ptr()), "e2e_fusion_merge_append_bwd")
            g_depth = None
            if ctx.needs_input_grad[0]:
                g_depth = torch.empty(H, W, **z)
                check(lib().e2e_rgbd_maps_bwd(ptr(depth), ptr(K), ptr(pose), H, W, ctypes.c_float(ctx.sigma), ptr(g_vg), None,
                                              ptr(g_alpha), ptr(g_depth), stream_ptr()), "e2e_rgbd_maps_bwd")
        return (g_depth, g_rgb if ctx.needs_input_grad[1] else None, None, None, go_pts, None, go_col, go_cc,
                None, None, None, None, None, None)


class PointFusion:
    """Point-based fusion SLAM (gradslam.slam.PointFusion as driven by the reference: constructor arguments
    at train_depth.py:111-118, .step at slam/custom_slam.py:33 and online_adaption.py:354-363, 466-469,
    __call__ at train_depth.py:266, 378-381)."""

    def __init__(self, odom="gradicp", dist_th=0.05, angle_th=20, sigma=0.6, dsratio=4, numiters=20, damp=1e-8,
                 dist_thresh=None, lambda_max=2.0, B=1.0, B2=1.0, nu=200.0, device=None):
        if odom not in ("gt", "icp", "gradicp"):
            raise ValueError(f"odometry method ({odom}) not supported for PointFusion")
        for name, v in (("dist_th", dist_th), ("angle_th", angle_th), ("sigma", sigma)):
            if not isinstance(v, (int, float)):
                raise TypeError(f"{name} must be a number. Got {type(v)}.")
        if dist_th < 0 or not (0 <= angle_th <= 90):
            raise ValueError("dist_th must be non-negative and angle_th within [0, 90]")
        self.odom, self.dist_th, self.angle_th, self.sigma = odom, float(dist_th), float(angle_th), float(sigma)
        self.dot_th = math.cos(angle_th * math.pi / 180.0)
        self.numiters, self.dsratio, self.damp, self.dist_thresh = int(numiters), int(dsratio), float(damp), dist_thresh
        self.lambda_max, self.B, self.B2, self.nu = float(lambda_max), float(B), float(B2), float(nu)
        self.device = torch.device(device) if device is not None else torch.device("cuda")
        self.last_association = None      # index_map / append_slot of the most recent step (tests, debugging)

    def _localize(self, pointclouds, live_frame, prev_frame):
        if self.odom == "gt" or prev_frame is None:
            if live_frame.poses is None:
                raise ValueError("live_frame.poses must be set when odom == 'gt' or prev_frame is None")
            return live_frame.poses
        from . import odometry
        if not isinstance(prev_frame, RGBDImages):
            raise TypeError(f"Expected prev_frame to be of type RGBDImages. Got {type(prev_frame)}.")
        if prev_frame.poses is None:
            raise ValueError("prev_frame.poses must be set for ICP odometry")
        B, _, H, W = live_frame.shape
        if len(pointclouds) != B or not pointclouds.has_points:
            raise ValueError("ICP odometry needs a non-empty map (fuse the first frame with prev_frame=None)")
        ds = self.dsratio
        poses = []
        for b in range(B):
            prev_pose = f32(prev_frame.poses[b, 0], "poses").detach().contiguous()
            depth = f32(live_frame.depth_image[b, 0, :, :, 0], "depth_image")
            K = f32(live_frame.intrinsics[b, 0], "intrinsics").contiguous()
            if torch.is_grad_enabled() and depth.requires_grad:       # differentiable vertex map (GradICP's purpose)
                us = torch.arange(W, dtype=torch.float32, device=depth.device)[None, :]
                vs = torch.arange(H, dtype=torch.float32, device=depth.device)[:, None]
                V = torch.stack([(us - K[0, 2]) / K[0, 0] * depth, (vs - K[1, 2]) / K[1, 1] * depth, depth], -1)
                vg, valid = V @ prev_pose[:3, :3].t() + prev_pose[:3, 3], depth > 0
            else:
                with torch.cuda.device(depth.device):
                    vg, _, _, valid = _frame_maps(depth.contiguous(), K, prev_pose, self.sigma)
                valid = valid.bool()
            live_pts = vg[::ds, ::ds][valid[::ds, ::ds]]
            m = pointclouds._maps[b]
            n = m.count()
            inside, h, w = odometry.active_map_points(m.pts[:n].detach(), K, prev_pose, H, W)
            keep = inside & (h % ds == 0) & (w % ds == 0)
            tgt, nrm = m.pts[:n].detach()[keep], m.nrm[:n].detach()[keep]
            if live_pts.shape[0] < 6 or tgt.shape[0] < 6:
                raise RuntimeError(f"ICP odometry: too few points (live {live_pts.shape[0]}, active map {tgt.shape[0]})")
            eye = torch.eye(4, dtype=torch.float32, device=depth.device)
            if self.odom == "icp":
                T, _ = odometry.point_to_plane_ICP(live_pts[None], tgt[None], nrm[None], eye, self.numiters, self.damp, self.dist_thresh)
            else:
                T, _ = odometry.point_to_plane_gradICP(live_pts[None], tgt[None], nrm[None], eye, self.numiters, self.damp,
                                                       self.dist_thresh, self.lambda_max, self.B, self.B2, self.nu)
            poses.append(T @ prev_pose)
        return torch.stack(poses).unsqueeze(1)

    def step(self, pointclouds, live_frame, prev_frame=None, inplace=False):
        if not isinstance(pointclouds, Pointclouds):
            raise TypeError(f"Expected pointclouds to be of type Pointclouds. Got {type(pointclouds)}.")
        if not isinstance(live_frame, RGBDImages):
            raise TypeError(f"Expected live_frame to be of type RGBDImages. Got {type(live_frame)}.")
        if live_frame.shape[1] != 1:
            raise ValueError(f"live_frame must have sequence length 1. Got {live_frame.shape[1]}.")
        poses = self._localize(pointclouds, live_frame, prev_frame)
        live_frame.poses = poses
        B, _, H, W = live_frame.shape
        if len(pointclouds) not in (0, B):
            raise ValueError(f"pointclouds batch size ({len(pointclouds)}) does not match the frame's ({B})")
        out = pointclouds if inplace else Pointclouds(device=pointclouds.device)
        maps = []
        self.last_association = []
        for b in range(B):
            old = pointclouds._maps[b] if len(pointclouds) else _Map(live_frame.device)
            depth = f32(live_frame.depth_image[b, 0, :, :, 0], "depth_image").contiguous()
            rgb = f32(live_frame.rgb_image[b, 0], "rgb_image").contiguous()
            K = f32(live_frame.intrinsics[b, 0], "intrinsics").contiguous()
            pose = f32(poses[b, 0], "poses").detach().contiguous()
            if old.n_upper > 4 * H * W + 2 * old.n_host:
                old.count()                     # occasionally tighten the bound (one sync) so buffers stay compact
            grad = torch.is_grad_enabled() and (depth.requires_grad or rgb.requires_grad or old.pts.requires_grad or
                                                old.col.requires_grad or old.cc.requires_grad)
            pts, nrm, col, cc, n_out, index_map, slot = _FusionStep.apply(
                depth, rgb, K, pose, old.pts, old.nrm, old.col, old.cc, old.n_dev, old.n_upper,
                self.dist_th, self.dot_th, self.sigma, bool(inplace and not grad))
            m = old if inplace else _Map(live_frame.device)
            m.pts, m.nrm, m.col, m.cc, m.n_dev = pts, nrm, col, cc, n_out
            m.n_upper, m.n_host = old.n_upper + H * W, old.n_host
            maps.append(m)
            self.last_association.append((index_map, slot))
        out._maps = maps
        return out, poses

    def _fuse_sequence(self, frames):
        """Known poses, no autograd: the frame loop runs inside the library -- e2e_fusion_sequence for one sequence, ONE
        cooperative launch over all B sequences (e2e_fusion_sequence_batch) for a batch -- instead of ~15 launches / memsets /
        allocations per frame and batch element from Python."""
        B, L, H, W = frames.shape[:4]
        dev = frames.device
        out = Pointclouds(device=dev)
        maps = []
        cap = L * H * W
        z = dict(dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            depth = f32(frames.depth_image[..., 0], "depth_image").contiguous()
            rgb = f32(frames.rgb_image, "rgb_image").contiguous()
            K = f32(frames.intrinsics[:, 0], "intrinsics").contiguous()
            poses = f32(frames.poses, "poses").contiguous()
            pts, nrm, col, cc = torch.empty(B, cap, 3, **z), torch.empty(B, cap, 3, **z), torch.empty(B, cap, 3, **z), torch.empty(B, cap, **z)
            n = torch.zeros(B, 2, dtype=torch.int64, device=dev)
            if B == 1:
                nws = lib().e2e_fusion_sequence_workspace_bytes(H, W, cap)
                ws = torch.empty(nws, dtype=torch.uint8, device=dev)
                check(lib().e2e_fusion_sequence(ptr(depth), ptr(rgb), ptr(K), ptr(poses), L, H, W, ctypes.c_float(self.sigma),
                                                ctypes.c_float(self.dist_th), ctypes.c_float(self.dot_th), ptr(pts), ptr(nrm),
                                                ptr(col), ptr(cc), ptr(n), 0, cap, ptr(ws), nws, stream_ptr()),
                      "e2e_fusion_sequence")
            else:
                nws = lib().e2e_fusion_sequence_batch_workspace_bytes(B, H, W, cap)
                ws = torch.empty(nws, dtype=torch.uint8, device=dev)
                check(lib().e2e_fusion_sequence_batch(ptr(depth), ptr(rgb), ptr(K), ptr(poses), B, L, H, W, ctypes.c_float(self.sigma),
                                                      ctypes.c_float(self.dist_th), ctypes.c_float(self.dot_th), ptr(pts), ptr(nrm),
                                                      ptr(col), ptr(cc), ptr(n), cap, ptr(ws), nws, stream_ptr()),
                      "e2e_fusion_sequence_batch")
        for b in range(B):
            m = _Map(dev)
            m.pts, m.nrm, m.col, m.cc, m.n_dev = pts[b], nrm[b], col[b], cc[b], n[b, :1]
            m.n_upper, m.n_host = cap, 0
            maps.append(m)
        out._maps = maps
        self.last_association = []
        return out

    def compact(self, pointclouds):
        """Shrink the map buffers to the current point count (one host synchronisation per batch element)."""
        for m in pointclouds._maps:
            n = m.count()
            m.pts, m.nrm, m.col, m.cc = m.pts[:n].clone(), m.nrm[:n].clone(), m.col[:n].clone(), m.cc[:n].clone()
        return pointclouds

    def forward(self, frames):
        if not isinstance(frames, RGBDImages):
            raise TypeError(f"Expected frames to be of type RGBDImages. Got {type(frames)}.")
        B, L = frames.shape[:2]
        no_grad = not (torch.is_grad_enabled() and (frames.depth_image.requires_grad or frames.rgb_image.requires_grad))
        if self.odom == "gt" and frames.poses is not None and no_grad and L > 0 and frames.device.type == "cuda":
            return self._fuse_sequence(frames), frames.poses
        pointclouds = Pointclouds(device=frames.device)
        prev, recovered = None, []
        for s in range(L):
            live = frames[:, s]
            if s == 0 and live.poses is None:
                live.poses = torch.eye(4, device=frames.device).view(1, 1, 4, 4).repeat(B, 1, 1, 1)
            pointclouds, pose = self.step(pointclouds, live, prev, inplace=True)
            prev = live if self.odom != "gt" else None
            recovered.append(pose)
        return pointclouds, torch.cat(recovered, 1)

    __call__ = forward


class ICPSLAM(PointFusion):
    """Placeholder for gradslam.slam.ICPSLAM (imported by the reference, selected only by MODEL.slam ==
    'ICPSLAM'): its map update is plain concatenation, which is out of this round's scope."""

    def step(self, *a, **k):
        raise NotImplementedError("ICPSLAM is out of scope; use PointFusion (configs/config.yaml:29)")


class _TransformPoints(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, transform):
        pts, T = points.contiguous(), transform.contiguous()
        out = torch.empty_like(pts)
        with torch.cuda.device(pts.device):
            check(lib().e2e_transform_points_fwd(ptr(pts), ptr(T), pts.shape[0], ptr(out), stream_ptr()), "e2e_transform_points_fwd")
        ctx.save_for_backward(pts, T)
        return out

    @staticmethod
    def backward(ctx, g):
        pts, T = ctx.saved_tensors
        g = f32(g, "grad").contiguous()
        gp = gT = None
        if ctx.needs_input_grad[0]:
            gp = torch.empty_like(pts)
            with torch.cuda.device(pts.device):
                check(lib().e2e_transform_points_bwd(ptr(T), ptr(g), pts.shape[0], ptr(gp), stream_ptr()), "e2e_transform_points_bwd")
        if ctx.needs_input_grad[1]:                           # 3x4 sums; the reference never asks for them on this path
            gT = torch.zeros_like(T)
            gT[:3, :3] = g.t() @ pts
            gT[:3, 3] = g.sum(0)
        return gp, gT


def transform_pointcloud(pointcloud, transform):
    """gradslam.geometry.geometryutils.transform_pointcloud: (N,3) points by a 4x4 rigid transform
    (online_adaption.py:642): one kernel each way (e2e_transform_points_*), the same left-to-right arithmetic as the query
    load that `losses.point_supervision_loss` fuses into the nearest-neighbour kernel."""
    if not torch.is_tensor(pointcloud) or not torch.is_tensor(transform):
        raise TypeError("pointcloud and transform must be tensors")
    if pointcloud.dim() != 2 or pointcloud.shape[1] != 3 or transform.shape != (4, 4):
        raise ValueError(f"expected pointcloud (N,3) and transform (4,4), got {tuple(pointcloud.shape)} / {tuple(transform.shape)}")
    f32(pointcloud, "pointcloud"), f32(transform, "transform")
    return _TransformPoints.apply(pointcloud, transform)


def find_active_map_points(pointclouds, rgbdimages):
    """gradslam.slam.fusionutils.find_active_map_points (imported by the reference at online_adaption.py:35):
    the map points that lie in front of the live camera and project into the frame, as `pc2im_bnhw` (N_active, 4) int64 rows
    (batch index, point index, pixel row, pixel column), ordered by (b, n).  `rgbdimages` is one live frame (sequence length 1)
    with its pose set.  One host synchronisation per batch element (the result's size)."""
    if not isinstance(pointclouds, Pointclouds):
        raise TypeError(f"Expected pointclouds to be of type Pointclouds. Got {type(pointclouds)}.")
    if not isinstance(rgbdimages, RGBDImages):
        raise TypeError(f"Expected rgbdimages to be of type RGBDImages. Got {type(rgbdimages)}.")
    if rgbdimages.shape[1] != 1:
        raise ValueError(f"Expected rgbdimages to have sequence length of 1. Got {rgbdimages.shape[1]}.")
    if rgbdimages.poses is None:
        raise ValueError("rgbdimages.poses must be set")
    B, _, H, W = rgbdimages.shape
    if len(pointclouds) != B:
        raise ValueError(f"Expected equal batch sizes for pointclouds and rgbdimages. Got {len(pointclouds)} and {B}.")
    dev = rgbdimages.device
    rows = []
    for b in range(B):
        m = pointclouds._maps[b]
        n = m.count()
        if n == 0:
            continue
        K = f32(rgbdimages.intrinsics[b, 0], "intrinsics").contiguous()
        pose = f32(rgbdimages.poses[b, 0], "poses").detach().contiguous()
        pts = m.pts.detach()
        out = torch.empty(n, 4, dtype=torch.int64, device=dev)
        n_act = torch.empty(1, dtype=torch.int64, device=dev)
        nws = lib().e2e_fusion_active_points_workspace_bytes(n)
        ws = torch.empty(nws, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib().e2e_fusion_active_points(ptr(pts), n, ptr(K), ptr(pose), H, W, b, ptr(out), ptr(n_act), ptr(ws), nws, stream_ptr()),
                  "e2e_fusion_active_points")
        rows.append(out[:int(n_act.item())])
    if not rows:
        return torch.zeros(0, 4, dtype=torch.int64, device=dev)
    return torch.cat(rows, 0)


def image_recover_slam(noisy_rgbd, slam, device):
    """slam/custom_slam.py:6-35: fuse a sequence frame by frame, detaching every frame but the last, so the
    gradient reaches only the last frame's depth and colour."""
    noisy_pointcloud = Pointclouds(device=device)
    batch_size, seq_len = noisy_rgbd.shape[:2]
    initial_poses = torch.eye(4, device=device).view(1, 1, 4, 4).repeat(batch_size, 1, 1, 1)
    for s in range(seq_len):
        live_frame = noisy_rgbd[:, s].to(device)
        live_frame = live_frame.detach() if s < seq_len - 1 else live_frame
        if s == 0 and live_frame.poses is None:
            live_frame.poses = initial_poses
        noisy_pointcloud, live_frame.poses = slam.step(noisy_pointcloud, live_frame)
        live_frame.poses = live_frame.poses.detach()
    return noisy_pointcloud
