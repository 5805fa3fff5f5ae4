"""A reduced set of hot-path invocations for compute-sanitizer (tools/sanitize.sh): the streaming kernel on tiny / ragged / TMA-eligible
shapes, the cooperative fusion sequence (single and batched), the grid kNN, and the round-2 entry points.  Small on purpose: the
sanitizer slows kernels by 10-100x."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "end-to-end-self-supervised-slam_b200"))
import e2e_slam_b200 as e2e  # noqa: E402
from e2e_slam_b200 import losses, ops  # noqa: E402
from e2e_slam_b200.slam import PointFusion, Pointclouds, RGBDImages, find_active_map_points, transform_pointcloud  # noqa: E402
from e2e_slam_b200.synthetic import make_pairs, room_sequence  # noqa: E402

dev = torch.device("cuda", 0)
for (B, H, W, pad) in [(3, 2, 2, "border"), (1, 3, 3, "zeros"), (1, 17, 5, "zeros"), (2, 48, 64, "border"), (1, 37, 52, "border"), (1, 61, 40, "zeros")]:
    d = {k: v.to(dev) for k, v in make_pairs(B, H, W, "tum", seed=H, rot_deg=4.0, trans=0.2).items()}
    depth = d["depth"].clone().requires_grad_(True)
    colors = d["colors"].clone().requires_grad_(True)
    T = d["T"].clone().requires_grad_(True)
    src, tgt = colors[:, 0].permute(0, 3, 1, 2), colors[:, 1].permute(0, 3, 1, 2)
    e2e.warp_photometric_loss(depth, d["inv_K"], d["K"], T, src, tgt, pad, True).backward()                    # lean sweep
    lm, syn, valid, pix = e2e.warp_photometric(depth, d["inv_K"], d["K"], T, src, tgt, pad, True, need_outputs=True)
    (lm * torch.rand_like(lm)).sum().backward()                                                                # map path, non-uniform upstream
    disp = (1.0 / d["depth"]).requires_grad_(True)
    ops.warp_photometric_loss_from_disparity(disp, d["inv_K"], d["K"], d["T"], src.detach(), tgt.detach(), torch.tensor(1.1, device=dev), pad, True).backward()
    e2e.warp_photometric_multi(depth, d["inv_K"], d["K"], [T, T], [src, src], tgt, pad, True, min_reprojection=True, auto_masking=True).backward()
    print("stream", B, H, W, pad, "ok")
# PointFusion: cooperative sequence kernel, single and batched, + per-frame steps
L, H, W = 4, 24, 32
depth, rgb, K, poses = room_sequence(L, H, W, device=dev)
one = RGBDImages(rgb[None], depth[None, ..., None], K.view(1, 1, 4, 4), poses[None])
two = RGBDImages(torch.stack([rgb, rgb.flip(2)]), torch.stack([depth, depth.flip(2)])[..., None], K.view(1, 1, 4, 4).repeat(2, 1, 1, 1), torch.stack([poses, poses]))
slam = PointFusion(odom="gt", device=dev)
with torch.no_grad():
    pc, _ = slam(one)
    pcb, _ = slam(two)
    pcs = Pointclouds(device=dev)
    for s in range(L):
        pcs, _ = slam.step(pcs, one[:, s], inplace=True)
    rows = find_active_map_points(pc, one[:, L - 1])
print("fusion ok", pc.points_list[0].shape, pcb.points_list[1].shape, rows.shape)
# kNN (grid and brute force), colour loss, transform, median
g = torch.Generator(device=dev).manual_seed(0)
ref = torch.rand(1, 70000, 3, generator=g, device=dev)
qry = (ref[:, :3000] + 0.01 * torch.rand(1, 3000, 3, generator=g, device=dev)).requires_grad_(True)
loss, idx = losses.knn_points_loss(ref, qry)                                    # 3000 x 70000 >= 2^24: grid path
loss.backward()
losses.knn_points_loss(ref[:, :500], qry[:, :200].detach())                     # brute force
col = torch.rand(1, 70000, 3, generator=g, device=dev)
losses.color_points_loss(col, col[:, :3000].clone().requires_grad_(True), idx).backward()
transform_pointcloud(qry[0], torch.eye(4, device=dev)).sum().backward()
ops.median(torch.rand(100001, generator=g, device=dev))
torch.cuda.synchronize()
print("points ok")
