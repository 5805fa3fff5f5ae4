"""BASELINE config C2: one synthetic TUM-shaped 480x640 key-frame pair (15 % depth holes, dilation-5 motion), refinement steps
on photometric + point supervision (weight 1.0, ~2 M-point global map) + depth smoothness (1e-3) + sparse depth supervision
(p = 0.012) -- the loss mix of train_depth.py:615-705 / online_adaption.py:638-645.  There is no depth network here (out of
scope, cuDNN): the predicted disparity itself is the optimised leaf (Adam, lr 1e-5 like configs/config.yaml).

A single pair is launch-latency bound: the figure of merit is time per refinement step, eager and as a CUDA
graph."""
import torch

from e2e_slam_b200 import losses, ops, view_synthesis
from e2e_slam_b200.synthetic import make_pairs


def _inputs(dev, H, W, map_points):
    g = torch.Generator(device=dev).manual_seed(3)
    d = make_pairs(1, H, W, "tum", seed=17, rot_deg=5.0, trans=0.15, device=dev)
    gt_depth = d["depth"].clone()
    gt_depth[torch.rand(1, 1, H, W, generator=g, device=dev) < 0.15] = 0.0
    disp0 = 1.0 / (d["depth"] * (1.0 + 0.05 * torch.randn(1, 1, H, W, generator=g, device=dev)))
    mask = ((torch.rand(1, H, W, 1, generator=g, device=dev) < 0.012) & (gt_depth.permute(0, 2, 3, 1) != 0)).float()
    sparse_gt = gt_depth.permute(0, 2, 3, 1) * mask
    # global map: the surface seen from the previous frame, densified to `map_points` with jitter
    reps = (map_points + H * W - 1) // (H * W)
    cam = view_synthesis.BackprojectDepth(1, H, W)(d["depth"], d["inv_K"])[0, :3].t()
    gmap = (cam.repeat(reps, 1)[:map_points] + 0.004 * torch.randn(map_points, 3, generator=g, device=dev)).contiguous()
    return d, disp0, sparse_gt, mask, gmap


def run(device, H=480, W=640, map_points=2_000_000, steps=3):
    dev = torch.device(device)
    d, disp0, sparse_gt, mask, gmap = _inputs(dev, H, W, map_points)
    src, tgt = d["colors"][:, 0].permute(0, 3, 1, 2), d["colors"][:, 1].permute(0, 3, 1, 2)
    bp = view_synthesis.BackprojectDepth(1, H, W)
    disp = disp0.clone().requires_grad_(True)
    opt = torch.optim.Adam([disp], lr=1e-5, capturable=True, fused=True)      # one optimizer kernel instead of torch's ~15 foreach launches
    terms = {}

    def step():
        opt.zero_grad(set_to_none=False)
        depth = 1.0 / disp
        photo = ops.warp_photometric_loss(depth, d["inv_K"], d["K"], d["T"], src, tgt, "border", True)
        smooth = losses.smoothness_loss(disp, tgt)
        gt_l1 = losses.depth_gt_loss(depth, sparse_gt, mask)
        cam = bp(depth, d["inv_K"])[0, :3].t()
        knn = losses.point_supervision_loss(cam, d["T"][0], gmap)
        loss = photo + 1.0 * knn + 1e-3 * smooth + gt_l1
        loss.backward()
        opt.step()
        terms.update(photo=photo.detach(), knn=knn.detach(), smooth=smooth.detach(), gt=gt_l1.detach(), loss=loss.detach())

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def timed(fn, n):
        a, b = ev(), ev()
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / n

    with torch.cuda.device(dev):
        l0 = ops.launch_count()
        step()
        per_step = ops.launch_count() - l0
        for _ in range(2):
            step()
        first = float(terms["loss"])
        eager_ms = timed(step, steps)
        graph_ms = None
        try:
            s = torch.cuda.Stream(device=dev)
            s.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s):
                step()
            torch.cuda.current_stream(dev).wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step()
            g.replay()
            graph_ms = timed(g.replay, max(steps, 10))
        except Exception as e:                    # capture is an optimisation, not part of the contract
            graph_ms = None
            terms["graph_error"] = str(e)[:200]
    out = {"workload": "C2: TUM-shaped 480x640 pair (15 % depth holes, dilation-5 motion), refinement step = photometric + point supervision "
                       f"(307 200 points vs {map_points} map points) + smoothness + sparse depth, backward, Adam on the disparity",
           "ms_per_step_eager": eager_ms, "ms_per_step_cuda_graph": graph_ms, "our_launches_per_step": per_step,
           "px_per_s_graph": None if graph_ms is None else H * W / (graph_ms * 1e-3), "px_per_s_eager": H * W / (eager_ms * 1e-3),
           "loss_terms": {k: float(v) for k, v in terms.items() if torch.is_tensor(v)}, "loss_after_warmup": first}
    if "graph_error" in terms:
        out["graph_error"] = terms["graph_error"]
    return out
