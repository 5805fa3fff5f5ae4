"""Full-size GPU parity cases (BASELINE configs at their real sizes): gradients of the fused warp + photometric kernel held to the
1e-5 contract away from the kinks of the loss, config C2 with the ~2 M-point map and all 307 200 live points, config C3's 60-frame
480x640 sequence bit for bit.  Slower than the rest of the suite (tens of seconds of CPU oracle work each)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_max, same_values

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _record(name, payload):
    """Measured errors go to stdout (pytest -s) and to gpurun_out/parity_measured.jsonl so that they can be quoted."""
    print(name, json.dumps(payload))
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity_measured.jsonl"), "a") as f:
            f.write(json.dumps({"case": name, **payload}) + "\n")
    except OSError:
        pass


def _dilate(mask, r):
    m = torch.from_numpy(mask.astype(np.float32))[None, None]
    return (torch.nn.functional.max_pool2d(m, 2 * r + 1, 1, r)[0, 0] > 0).numpy()


def _kink_masks(d, pad, mask):
    """Target pixels at which the reference's fp32 and fp64 evaluations take DIFFERENT branches of a non-smooth operation, so that
    their gradients differ by O(1) and neither is "the" gradient: floor() of the sampling coordinate, the border clamp, the validity
    test |grid| <= 1, the SSIM clamp to [0,1], and sign(x - y) of the L1 term."""
    from oracle import torch_oracle as to
    out = {}
    for dt in (torch.float32, torch.float64):
        c = lambda t: t.to(dt)
        src, tgt = c(d["colors"][:, 0]).permute(0, 3, 1, 2), c(d["colors"][:, 1]).permute(0, 3, 1, 2)
        B, _, H, W = d["depth"].shape
        pix, valid = to.project(to.backproject(c(d["depth"]), c(d["inv_K"])), c(d["K"]), c(d["T"]), H, W)
        ix, iy = ((pix[..., 0] + 1) * W - 1) / 2, ((pix[..., 1] + 1) * H - 1) / 2
        inb = (ix > 0) & (ix < W - 1) & (iy > 0) & (iy < H - 1)
        if pad == "border":
            ix, iy = ix.clamp(0, W - 1), iy.clamp(0, H - 1)
        syn = torch.nn.functional.grid_sample(src, pix, padding_mode=pad, align_corners=False)
        x, y = (syn * valid, tgt * valid) if mask else (syn, tgt)
        raw = to.ssim(x, y, return_raw=True)
        out[dt] = dict(x0=torch.floor(ix), y0=torch.floor(iy), inb=inb, valid=valid[:, 0] > 0, lo=(raw < 0).any(1), hi=(raw > 1).any(1),
                       sg=torch.sign(x - y))
    a, b = out[torch.float32], out[torch.float64]
    local = (a["x0"] != b["x0"].float()) | (a["y0"] != b["y0"].float()) | (a["inb"] != b["inb"]) | (a["sg"] != b["sg"].float()).any(1)
    wide = (a["valid"] != b["valid"]) | (a["lo"] != b["lo"]) | (a["hi"] != b["hi"])
    return local.numpy(), wide.numpy(), a["x0"].long().numpy(), a["y0"].long().numpy()


@pytest.mark.parametrize("H,W,kind,rot,trans", [(480, 640, "icl", 2.0, 0.05), (480, 640, "tum", 5.0, 0.15), (1080, 1920, "icl", 2.0, 0.05)])
def test_gradients_hold_1e5_away_from_kinks(H, W, kind, rot, trans):
    """north_star: gradients within 1e-5 relative (fp32).  At BASELINE sizes the reference's own fp32 gradient differs from its float64
    evaluation by 0.25 .. 0.77 of max|g| in the max norm -- at the handful of pixels where the two precisions take different branches
    of floor / clamp / |.| / sign -- and by 1.4e-4 .. 2.6e-4 everywhere else (fp32 cancellation in its `grad_c . q` chain).  The kink
    pixels (and the pixels / source texels their windows and taps reach) are excluded; EVERYWHERE ELSE the kernel's gradient must
      * be within 1e-5 (relative to max |ref|) of the reference's fp32 gradient at 99.9 % of the pixels, and
      * nowhere be further from it than max(1e-5, a tenth of the reference's own fp32-vs-float64 deviation there)
    (measured on B200: 5.7e-6 .. 1.8e-5 worst pixel for grad_depth, <= 9.8e-6 for grad_src; the numbers are recorded)."""
    import e2e_slam_b200 as e2e
    from e2e_slam_b200.synthetic import make_pairs
    from oracle import torch_oracle as to
    d = make_pairs(1, H, W, kind, seed=H + W, rot_deg=rot, trans=trans)
    depth = d["depth"].cuda().requires_grad_(True)
    colors = d["colors"].cuda().requires_grad_(True)
    src, tgt = colors[:, 0].permute(0, 3, 1, 2), colors[:, 1].permute(0, 3, 1, 2)
    e2e.warp_photometric_loss(depth, d["inv_K"].cuda(), d["K"].cuda(), d["T"].cuda(), src, tgt, "border", True).backward()
    r32 = to.fwd_bwd(d["depth"], d["inv_K"], d["K"], d["T"], d["colors"][:, 0], d["colors"][:, 1], "border", True, want=("depth", "src"))
    r64 = to.fwd_bwd(d["depth"], d["inv_K"], d["K"], d["T"], d["colors"][:, 0], d["colors"][:, 1], "border", True, dtype=torch.float64,
                     want=("depth", "src"))
    local, wide, x0, y0 = _kink_masks(d, "border", True)
    kink = _dilate(local[0], 1) | _dilate(wide[0], 3)                      # target pixels whose gradient a kink can reach
    # source texels touched by the taps of kinked target pixels
    ys, xs = np.nonzero(kink)
    tex = np.zeros((H, W), bool)
    for dy in (0, 1):
        for dx in (0, 1):
            yy, xx = np.clip(y0[0][ys, xs] + dy, 0, H - 1), np.clip(x0[0][ys, xs] + dx, 0, W - 1)
            tex[yy, xx] = True
    rep = {"shape": [H, W], "kink_px": int(kink.sum()), "kink_texels": int(tex.sum())}
    for name, ours, a, b, keep in (("g_depth", depth.grad[0, 0].cpu().numpy(), r32["g_depth"][0, 0].numpy(), r64["g_depth"][0, 0].numpy(), ~kink),
                                   ("g_src", colors.grad[0, 0].cpu().numpy(), r32["g_src"][0].numpy(), r64["g_src"][0].numpy(), ~tex)):
        scale = float(np.abs(b).max())
        e32, e64 = np.abs(ours - a) / scale, np.abs(ours - b) / scale
        best = np.minimum(e32, e64)
        rep[name] = {"all_px_vs_ref32": float(e32.max()), "all_px_vs_ref64": float(e64.max()), "ref32_vs_ref64_all_px": float((np.abs(a - b) / scale).max()),
                     "away_from_kinks_vs_ref32": float(e32[keep].max()), "away_from_kinks_vs_ref64": float(e64[keep].max()),
                     "away_from_kinks_p999_vs_ref32": float(np.quantile(e32[keep], 0.999)),
                     "away_from_kinks_best": float(best[keep].max()), "ref32_vs_ref64_away_from_kinks": float((np.abs(a - b) / scale)[keep].max())}
    _record(f"grad_kinks_{kind}_{H}x{W}", rep)
    assert rep["kink_px"] < 0.02 * H * W                                   # the exclusion really is a handful of pixels
    for name in ("g_depth", "g_src"):
        r = rep[name]
        assert r["away_from_kinks_p999_vs_ref32"] <= RTOL, (name, r)
        assert r["away_from_kinks_vs_ref32"] <= max(RTOL, 0.1 * r["ref32_vs_ref64_away_from_kinks"]), (name, r)


def test_c2_full_size_step():
    """Config C2 at its real size: one refinement step's loss mix (photometric + point supervision + smoothness + sparse depth) with a
    2.1 M-point global map and ALL 307 200 live points; the CPU side finds the neighbours with scipy's cKDTree and evaluates the
    reference composition in float64 from the same fp32 inputs.  Loss terms and the disparity gradient to 1e-5."""
    from scipy.spatial import cKDTree
    import e2e_slam_b200 as e2e
    from e2e_slam_b200 import losses, view_synthesis
    from e2e_slam_b200.synthetic import make_pairs
    from oracle import torch_oracle as to
    from test_configs_gpu import _c2_inputs
    d, disp0, sparse_gt, mask, _ = _c2_inputs()
    g = torch.Generator().manual_seed(21)
    # ~2.1 M map points: the true surface sampled at 7 jittered sub-pixel positions per pixel, in the previous frame
    K, depth_gt = d["K"][0], d["depth"][0, 0]
    ys, xs = torch.meshgrid(torch.arange(480.0), torch.arange(640.0), indexing="ij")
    pts = []
    for _ in range(7):
        jx, jy = xs + torch.rand(480, 640, generator=g) - 0.5, ys + torch.rand(480, 640, generator=g) - 0.5
        z = depth_gt * (1.0 + 0.003 * torch.randn(480, 640, generator=g))
        pts.append(torch.stack([(jx - K[0, 2]) / K[0, 0] * z, (jy - K[1, 2]) / K[1, 1] * z, z], -1).reshape(-1, 3))
    gmap = torch.cat(pts).contiguous()
    assert gmap.shape[0] == 7 * 480 * 640

    def oracle(dtype):
        c = lambda t: t.to(dtype)
        disp = c(disp0).clone().requires_grad_(True)
        depth = 1.0 / disp
        src, tgt = c(d["colors"][:, 0]).permute(0, 3, 1, 2), c(d["colors"][:, 1]).permute(0, 3, 1, 2)
        photo = to.warp_photometric(depth, c(d["inv_K"]), c(d["K"]), c(d["T"]), src, tgt, "border", True)[0].mean()
        smooth = to.smoothness(disp, tgt)
        gt_l1 = to.sparse_gt_l1(depth, c(sparse_gt), c(mask))
        cam = to.backproject(depth, c(d["inv_K"]))[0, :3].t()
        q = cam @ c(d["T"])[0, :3, :3].t() + c(d["T"])[0, :3, 3]
        _, idx = cKDTree(gmap.numpy().astype(np.float64)).query(q.detach().numpy().astype(np.float64), k=1, workers=-1)
        knn = ((q - c(gmap)[torch.from_numpy(idx)]) ** 2).sum(1).mean()
        (photo + 1.0 * knn + 1e-3 * smooth + gt_l1).backward()
        return dict(photo=float(photo), knn=float(knn), smooth=float(smooth), gt=float(gt_l1)), disp.grad.double().numpy()

    t32, g32 = oracle(torch.float32)
    t64, g64 = oracle(torch.float64)
    cu = {k: v.cuda() for k, v in d.items()}
    disp = disp0.cuda().requires_grad_(True)
    depth = 1.0 / disp
    src, tgt = cu["colors"][:, 0].permute(0, 3, 1, 2), cu["colors"][:, 1].permute(0, 3, 1, 2)
    photo = e2e.warp_photometric_loss(depth, cu["inv_K"], cu["K"], cu["T"], src, tgt, "border", True)
    smooth = losses.smoothness_loss(disp, tgt)
    gt_l1 = losses.depth_gt_loss(depth, sparse_gt.cuda(), mask.cuda())
    cam = view_synthesis.BackprojectDepth(1, 480, 640)(depth, cu["inv_K"])[0, :3].t()
    knn = losses.point_supervision_loss(cam, cu["T"][0], gmap.cuda())
    (photo + 1.0 * knn + 1e-3 * smooth + gt_l1).backward()
    ours = dict(photo=float(photo), knn=float(knn), smooth=float(smooth), gt=float(gt_l1))
    go = disp.grad.double().cpu().numpy()
    l2 = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    rep = {"terms_ours": ours, "terms_ref64": t64, "grad_l2_vs_ref32": l2(go, g32), "grad_l2_vs_ref64": l2(go, g64), "ref32_vs_ref64_l2": l2(g32, g64),
           "grad_max_vs_ref64": rel_max(go, g64), "ref32_vs_ref64_max": rel_max(g32, g64), "map_points": int(gmap.shape[0]), "queries": 480 * 640}
    _record("c2_full_size", rep)
    for k in t64:
        assert abs(ours[k] - t64[k]) <= RTOL * max(abs(t64[k]), 1e-6), (k, ours[k], t64[k])
    assert min(rep["grad_l2_vs_ref64"], rep["grad_l2_vs_ref32"]) <= max(RTOL, rep["ref32_vs_ref64_l2"]), rep


def test_c3_sixty_frames_bit_exact():
    """Config C3 as benchmarked: 60 frames x 480x640 through ONE e2e_fusion_sequence launch (PointFusion.__call__), against the numpy
    oracle fusing the same frames one by one: final map (1.3 M points), every attribute, bit for bit -- the sequence kernel's 128-bit
    CAS races, look-back appends and grid barriers at the size where they are stressed."""
    from e2e_slam_b200.slam import PointFusion, RGBDImages
    from oracle import fusion_oracle as fo
    L, H, W = 60, 480, 640
    depth, rgb, K, poses = fo.synthetic_room_sequence(L, H, W, seed=0)
    depth, rgb = depth.astype(np.float32), rgb.astype(np.float32)
    depth[:, 100:140, 200:260] = 0.0
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    frames = RGBDImages(t(rgb)[None], t(depth)[None, ..., None], t(K).view(1, 1, 4, 4), t(poses)[None])
    with torch.no_grad():
        pc, _ = PointFusion(odom="gt", dist_th=0.05, angle_th=20, sigma=0.6, device="cuda")(frames)
        got = [x.cpu().numpy() for x in (pc.points_list[0], pc.normals_list[0], pc.colors_list[0], pc.features_list[0][:, 0])]
    o = fo.PointFusionOracle(0.05, 20, 0.6)
    for s in range(L):
        o.step(depth[s], rgb[s], K, poses[s])
    _record("c3_60_frames", {"map_points": int(len(o.points)), "live_points": int((depth > 0).sum())})
    assert got[0].shape[0] == len(o.points)
    for name, a, b in zip(("points", "normals", "colors", "ccount"), got, (o.points, o.normals, o.colors, o.ccount)):
        assert same_values(a, b) == 0, f"{name}: {same_values(a, b)} of {a.size} values differ"
