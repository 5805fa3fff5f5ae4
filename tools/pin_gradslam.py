#!/usr/bin/env python
"""Regenerate golden vectors for everything gradslam / chamferdist define, from the REAL packages, the moment they are importable.

    python tools/pin_gradslam.py [--out tests/golden_gradslam]

The reference imports gradslam (PointFusion, RGBDImages, Pointclouds, find_active_map_points, transform_pointcloud, the ICP
odometry providers) and chamferdist (knn_points) but vendors neither, pins no version, and ships no test or fixture for them
(SURVEY.md section 8(c)); neither is installed in the build container and there is no network, so oracle/fusion_oracle.py,
oracle/fusion_oracle_torch.py and oracle/icp_oracle.py are restatements and their parity is UNPINNED.  This script is the pin:
run it in any environment where `import gradslam` / `import chamferdist` succeed (pip install gradslam==0.1.0; chamferdist at the
commit loss/losses.py:42 cites) and commit the .npz files it writes.  tests/test_gradslam_pin.py picks them up automatically
(it is skipped with the reason "parity unpinned" while the directory is empty) and checks the numpy oracle on CPU and the CUDA
kernels on the GPU against them -- integer outputs bit for bit, floats to 1e-5.

Inputs are the same seeded synthetic sequences the tests use (oracle/fusion_oracle.synthetic_room_sequence), so the goldens are
small (a 6-frame 60x80 sequence, a 19 200 x 75 000 ICP pair thinned to 2 000 x 5 000, a kNN case with exact ties).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden_gradslam"))
    args = ap.parse_args()
    missing = []
    try:
        import gradslam
        from gradslam import Pointclouds, RGBDImages
        from gradslam.slam import PointFusion
        from gradslam.slam.fusionutils import find_active_map_points
        from gradslam.geometry.geometryutils import transform_pointcloud
        from gradslam.odometry.icputils import point_to_plane_ICP, point_to_plane_gradICP
    except Exception as e:                                   # noqa: BLE001 -- any import problem means "not available here"
        missing.append(f"gradslam ({type(e).__name__}: {e})")
    try:
        from chamferdist.chamfer import knn_points
    except Exception as e:                                   # noqa: BLE001
        missing.append(f"chamferdist ({type(e).__name__}: {e})")
    if missing:
        print("cannot pin: " + "; ".join(missing))
        print("nothing written; oracle/fusion_oracle*.py and oracle/icp_oracle.py stay 'parity unpinned'")
        return 2
    import torch
    from oracle import fusion_oracle as fo
    os.makedirs(args.out, exist_ok=True)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))

    # ---- PointFusion(odom='gt') frame by frame: derived maps, active rows, map after every step --------------------------
    L, H, W = 6, 60, 80
    depth, rgb, K, poses = fo.synthetic_room_sequence(L, H, W, seed=3)
    depth = (depth * (np.random.default_rng(4).random(depth.shape) >= 0.1)).astype(np.float32)
    rgb = rgb.astype(np.float32)
    frames = RGBDImages(t(rgb)[None], t(depth)[None, ..., None], t(K).view(1, 1, 4, 4), t(poses)[None])
    slam = PointFusion(odom="gt", dist_th=0.05, angle_th=20, sigma=0.6, device="cpu")
    pc = Pointclouds(device="cpu")
    out = dict(depth=depth, rgb=rgb, K=K, poses=poses, gradslam_version=str(getattr(gradslam, "__version__", "unknown")))
    for s in range(L):
        live = frames[:, s]
        if pc.has_points:
            out[f"active_rows_{s}"] = find_active_map_points(pc, live).numpy()
        out[f"vertex_g_{s}"] = live.global_vertex_map[0, 0].numpy()
        out[f"normal_g_{s}"] = live.global_normal_map[0, 0].numpy()
        out[f"valid_{s}"] = live.valid_depth_mask[0, 0, ..., 0].numpy()
        pc, _ = slam.step(pc, live)
        out[f"points_{s}"] = pc.points_list[0].numpy()
        out[f"normals_{s}"] = pc.normals_list[0].numpy()
        out[f"colors_{s}"] = pc.colors_list[0].numpy()
        out[f"ccount_{s}"] = pc.features_list[0][:, 0].numpy()
    np.savez_compressed(os.path.join(args.out, "pointfusion_room_6x60x80.npz"), **out)

    # ---- transform_pointcloud ------------------------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(1)
    p = torch.randn(5001, 3, generator=g) * 2
    T = t(poses[3])
    np.savez_compressed(os.path.join(args.out, "transform_pointcloud.npz"), points=p.numpy(), T=T.numpy(),
                        out=transform_pointcloud(p, T).numpy())

    # ---- ICP / GradICP on a thinned room pair ------------------------------------------------------------------------------
    m0 = fo.rgbd_maps(depth[0], rgb[0], K, poses[0], 0.6)
    m1 = fo.rgbd_maps(depth[1], rgb[1], K, poses[0], 0.6)         # live frame placed with the PREVIOUS pose
    tgt, nrm = m0["vertex_g"][m0["valid"]], m0["normal_g"][m0["valid"]]
    src = m1["vertex_g"][::2, ::2][m1["valid"][::2, ::2]]
    eye = torch.eye(4)
    T_icp, idx_icp = point_to_plane_ICP(t(src)[None], t(tgt)[None], t(nrm)[None], eye, numiters=20, damp=1e-8, dist_thresh=None)
    T_g, idx_g = point_to_plane_gradICP(t(src)[None], t(tgt)[None], t(nrm)[None], eye, numiters=20, damp=1e-8, dist_thresh=None,
                                        lambda_max=2.0, B=1.0, B2=1.0, nu=200.0)
    np.savez_compressed(os.path.join(args.out, "icp_room_pair.npz"), src=src, tgt=tgt, nrm=nrm, T_icp=T_icp.numpy(),
                        idx_icp=idx_icp.numpy(), T_gradicp=T_g.numpy(), idx_gradicp=idx_g.numpy())

    # ---- chamferdist knn_points, K = 1, with exact ties ---------------------------------------------------------------------
    rng = np.random.default_rng(7)
    ref = rng.random((3000, 3)).astype(np.float32)
    ref[1500:1600] = ref[100:200]                                # duplicated reference points: which index wins a tie?
    qry = np.concatenate([ref[100:150], rng.random((2000, 3)).astype(np.float32)])
    kn = knn_points(t(qry)[None], t(ref)[None], K=1)
    np.savez_compressed(os.path.join(args.out, "knn_ties.npz"), query=qry, ref=ref, dists=kn.dists[0, :, 0].numpy(),
                        idx=kn.idx[0, :, 0].numpy())
    print(f"wrote goldens to {args.out}; now run: python -m pytest tests/test_gradslam_pin.py -q")
    return 0


if __name__ == "__main__":
    sys.exit(main())
