"""Drop-in wiring for the reference's driver scripts (train_depth.py, online_adaption.py, ...).

Tier (i) -- unmodified scripts, one line added before their imports:

    import e2e_slam_b200.patch as e2e_patch; e2e_patch.install()

`install()` pre-seeds `sys.modules` so that the scripts' own import statements
    from loss.losses import *                                   (train_depth.py:22)
    from depth_estimation.view_synthesis import BackprojectDepth, Project3D          (:30)
    from slam.custom_slam import image_recover_slam              (:27)
    from gradslam.slam import PointFusion, ICPSLAM ; from gradslam import Pointclouds, RGBDImages   (:33-38)
    from gradslam.geometry.geometryutils import transform_pointcloud                 (online_adaption.py:36)
    from chamferdist import ChamferDistance ; from chamferdist.chamfer import knn_points (loss/losses.py:3)
resolve to this package.  `depth_estimation.networks`, `utils.*` and the dataset classes stay the
reference's / gradslam's own (they are host code outside the hot path).

Tier (ii) -- the fused op behind the scripts' method names:

    e2e_patch.fuse(Depth_Estimation)        # or SLAM from online_adaption.py

replaces `novel_view_synthesis` and `compute_photometric_loss` (train_depth.py:545-613, 707-727) by versions
that run ONE fused forward kernel per source frame and hand autograd ONE fused backward kernel, while still
filling `outputs[("synthesized_frame", f)]` / `outputs[("valid_mask", f)]` that the scripts' plotting reads.
"""
import sys
import types

import torch

from . import losses, odometry, ops, slam, view_synthesis


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__e2e_slam_b200__ = True
    return m


def install(grid_sample=False, keep_real_gradslam_datasets=True):
    """Route the reference's imports to e2e_slam_b200.  Idempotent.  With grid_sample=True,
    torch.nn.functional.grid_sample is also replaced for CUDA fp32 bilinear zeros|border calls."""
    sm = sys.modules
    # --- the reference's own modules on the hot path -----------------------------------------------
    public = {k: v for k, v in vars(losses).items() if not k.startswith("_")}
    sm["loss.losses"] = _module("loss.losses", **public, __all__=[k for k in public if k not in ("torch", "nn", "ops", "namedtuple")])
    sm["depth_estimation.view_synthesis"] = view_synthesis
    sm["slam.custom_slam"] = _module("slam.custom_slam", image_recover_slam=slam.image_recover_slam, Pointclouds=slam.Pointclouds)
    # --- gradslam -------------------------------------------------------------------------------------
    real_datasets = None
    if keep_real_gradslam_datasets and "gradslam" not in sm:
        try:
            import importlib
            real_datasets = importlib.import_module("gradslam.datasets")
        except Exception:
            real_datasets = None
        for k in [k for k in sm if k == "gradslam" or k.startswith("gradslam.")]:
            if k != "gradslam.datasets" and not k.startswith("gradslam.datasets."):
                del sm[k]

    def _no_dataset(*a, **k):
        raise ImportError("gradslam.datasets (ICL/TUM loaders) is dataset IO outside the hot path; install gradslam "
                          "to use it, or feed tensors of the same layout (see e2e_slam_b200.synthetic)")

    fusionutils = _module("gradslam.slam.fusionutils", find_active_map_points=slam.find_active_map_points)
    gs_slam = _module("gradslam.slam", PointFusion=slam.PointFusion, ICPSLAM=slam.ICPSLAM, fusionutils=fusionutils)
    geomutils = _module("gradslam.geometry.geometryutils", transform_pointcloud=slam.transform_pointcloud)
    se3utils = _module("gradslam.geometry.se3utils", se3_exp=odometry.se3_exp)
    geometry = _module("gradslam.geometry", geometryutils=geomutils, se3utils=se3utils)
    icputils = _module("gradslam.odometry.icputils", point_to_plane_ICP=odometry.point_to_plane_ICP,
                       point_to_plane_gradICP=odometry.point_to_plane_gradICP, gauss_newton_solve=odometry.gauss_newton_solve,
                       solve_linear_system=odometry.solve_linear_system)
    gs_odometry = _module("gradslam.odometry", icputils=icputils)
    structures = _module("gradslam.structures", Pointclouds=slam.Pointclouds, RGBDImages=slam.RGBDImages)
    datasets = real_datasets or _module("gradslam.datasets", ICL=_no_dataset, TUM=_no_dataset)
    gs = _module("gradslam", Pointclouds=slam.Pointclouds, RGBDImages=slam.RGBDImages, slam=gs_slam, geometry=geometry,
                 structures=structures, datasets=datasets, odometry=gs_odometry)
    gs.__path__ = []
    sm.update({"gradslam": gs, "gradslam.slam": gs_slam, "gradslam.slam.fusionutils": fusionutils, "gradslam.geometry": geometry,
               "gradslam.geometry.geometryutils": geomutils, "gradslam.geometry.se3utils": se3utils, "gradslam.odometry": gs_odometry,
               "gradslam.odometry.icputils": icputils, "gradslam.structures": structures, "gradslam.datasets": datasets})
    # --- chamferdist ----------------------------------------------------------------------------------
    chamfer = _module("chamferdist.chamfer", knn_points=losses.knn_points)
    cd = _module("chamferdist", ChamferDistance=losses.ChamferDistance, chamfer=chamfer)
    cd.__path__ = []
    sm.update({"chamferdist": cd, "chamferdist.chamfer": chamfer})
    if grid_sample:
        _patch_grid_sample()


_torch_grid_sample = torch.nn.functional.grid_sample


def _patch_grid_sample():
    def grid_sample(input, grid, mode="bilinear", padding_mode="zeros", align_corners=None):
        if (input.is_cuda and input.dtype == torch.float32 and input.dim() == 4 and mode == "bilinear"
                and padding_mode in ("zeros", "border")):
            return view_synthesis.grid_sample(input, grid, mode, padding_mode, bool(align_corners))
        return _torch_grid_sample(input, grid, mode=mode, padding_mode=padding_mode, align_corners=align_corners)
    torch.nn.functional.grid_sample = grid_sample


# ---- tier (ii): fused replacements for the scripts' methods ------------------------------------------
def fused_novel_view_synthesis(self, inputs):
    """Replacement for Depth_Estimation.novel_view_synthesis / SLAM.novel_view_synthesis (non-geometric
    branch, train_depth.py:578-590): one fused kernel per source frame."""
    if self.args.LOSS.geometric:
        raise NotImplementedError("the fused op covers the photometric branch; LOSS.geometric uses the granular ops")
    outputs = {}
    # min-reprojection / auto-masking send a per-pixel gradient back (the winner of each pixel): no point in the speculative
    # uniform-gradient sweep then -- forward kernel now, gradient-map sweep in backward
    uniform = not (getattr(self.args.LOSS, "min_reprojection", False) or getattr(self.args.LOSS, "auto_masking", False))
    for frame in self.args.DATA.frames[1:]:
        T = inputs["T", frame]
        if T.dim() == 4:
            T = T.squeeze(1)                      # online_adaption.py:446
        loss_map, syn, valid, _ = ops.warp_photometric(
            inputs["target_depth"], inputs["Inverse_K"], inputs["K"], T, inputs["source_frame", frame], inputs["target_frame"],
            padding_mode=self.args.MODEL.padding_mode, photometric_mask=bool(self.args.LOSS.photometric_mask), need_outputs=True,
            expect_uniform=uniform)
        outputs[("valid_mask", frame)] = valid
        outputs[("synthesized_frame", frame)] = syn
        outputs[("photometric_map", frame)] = loss_map
    return outputs


def fused_compute_photometric_loss(self, inputs, outputs):
    """Replacement for compute_photometric_loss (train_depth.py:707-727): the per-frame maps were produced by
    the fused forward already; concatenate them exactly as the reference does (:726)."""
    return torch.cat([outputs[("photometric_map", frame)] for frame in self.args.DATA.frames[1:]], 1)


def fused_compute_losses(self, inputs, outputs):
    """Replacement for compute_losses (train_depth.py:615-705; online_adaption.py's SLAM.compute_losses is the same sequence): the same
    terms in the same order and weights, but
      * the frame reduction (mean / min-reprojection / auto-masking, :621-660) goes through losses.photometric_objective (one kernel
        for the per-pixel minimum instead of torch.cat + torch.min), and
      * no term is read back on its own: the reference synchronises the host once PER TERM (`.item()` at :662, 668, 673, 678, 684, 692,
        697) and once more for the return value; here every term stays a device scalar and ONE stacked read at the end serves all of
        them (`self.last_losses` holds the floats afterwards).  SURVEY.md 8(f) rank 3."""
    terms = {}
    self.optimizer.zero_grad()
    maps = self.compute_photometric_loss(inputs=inputs, outputs=outputs)
    frames = self.args.DATA.frames[1:]
    photo_maps = [maps[:, i:i + 1] for i in range(maps.shape[1])]
    ident = None
    if self.args.LOSS.auto_masking:
        am = self.compute_automasking_loss(inputs=inputs, outputs=outputs)
        ident = [am[:, i:i + 1] for i in range(len(frames))]
    optimize = losses.photometric_objective(photo_maps, ident, bool(self.args.LOSS.min_reprojection))
    loss = optimize
    terms["photometric_loss"] = optimize
    if self.args.LOSS.geometric:
        geometric = self.compute_geometric_loss(outputs=outputs).mean()
        loss = loss + geometric * self.args.LOSS.geometric_weight
        terms["geometric_loss"] = geometric
    if self.args.LOSS.smoothness:
        smooth_loss = self.compute_smoothness_loss(inputs=inputs)
        loss = loss + smooth_loss * self.args.LOSS.smoothness_weight
        terms["smoothn_loss"] = smooth_loss
    if self.args.LOSS.depth_regularizer:
        depth_reg = self.compute_depth_regularizer(inputs=inputs)
        loss = loss + depth_reg * self.args.LOSS.depth_regularizer_weight
        terms["depth regularizer"] = depth_reg
    if self.args.LOSS.knn_points:
        knn_loss, _ = losses.knn_points_loss(gt_pointcloud=self.gt_reconstruction.points_list[0].unsqueeze(0).contiguous(),
                                             noisy_pointcloud=inputs["noisy_pointcloud"])
        loss = loss + knn_loss * self.args.LOSS.knn_points_weight
        terms["knn_loss"] = knn_loss
    if self.args.LOSS.chamfer_distance:
        chamfer_dist = 0.5 * self.chamfer(inputs["noisy_pointcloud"], self.gt_reconstruction.points_list[0].unsqueeze(0).contiguous(),
                                          bidirectional=True)
        loss = loss + chamfer_dist * self.args.LOSS.chamfer_weight
        terms["chamfer_loss"] = chamfer_dist
    if self.args.LOSS.supervise_depth:
        gt_loss = self.compute_gt_depth_loss(inputs=inputs)
        loss = loss + gt_loss * self.args.LOSS.gt_depth_weight
        terms["gt_depth_loss"] = gt_loss
    loss.backward()
    self.optimizer.step()
    names = list(terms)
    values = torch.stack([terms[k].detach().reshape(()) for k in names] + [loss.detach().reshape(())]).tolist()     # the one host read
    self.last_losses = dict(zip(names, values[:-1]))
    return values[-1]


class GraphedStep:
    """A whole refinement step -- forward through the fused ops, loss terms, backward, optimizer step -- captured ONCE as a CUDA graph
    and replayed (a single key-frame pair is launch-bound, ~25 launches and several allocations per step).

        step = GraphedStep(fn)            # fn() runs one step on tensors it closes over and returns a dict of device scalars
        terms = step()                    # first calls: warm-up (eager, on a side stream) then capture; later calls: graph replay
    `fn` must be capture-safe: no host synchronisation (.item(), points_list, ...), static shapes, optimizer built with
    capturable=True, `zero_grad(set_to_none=False)`.  Inputs change between replays by copying into the tensors `fn` closes over.
    Every op of this package launches on the current stream and allocates only through torch, so it is capture-safe."""

    def __init__(self, fn, warmup=2):
        self.fn, self.warmup = fn, warmup
        self.graph, self.out, self.calls = None, None, 0

    def __call__(self):
        if self.graph is not None:
            self.graph.replay()
            return self.out
        dev = torch.cuda.current_device()
        if self.calls < self.warmup:                        # eager warm-up on a side stream (lazy initialisation, autograd buffers)
            s = torch.cuda.Stream(device=dev)
            s.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s):
                out = self.fn()
            torch.cuda.current_stream(dev).wait_stream(s)
            self.calls += 1
            return out
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.out = self.fn()
        self.graph = g
        g.replay()                                          # capture records, it does not execute
        return self.out


def fuse(cls, defer_items=True, graph=False):
    """Swap a reference driver class's view-synthesis + photometric-loss methods for the fused op.
    defer_items: also replace `compute_losses` by the version with ONE host read per step instead of one per loss term.
    graph:       give the class a `graphed_step(fn)` factory (GraphedStep) for whole-step CUDA-graph capture; the capture itself is
                 explicit because the step must then be free of host synchronisation (see GraphedStep)."""
    cls.novel_view_synthesis = fused_novel_view_synthesis
    cls.compute_photometric_loss = fused_compute_photometric_loss
    if defer_items:
        cls.compute_losses = fused_compute_losses
    if graph:
        cls.graphed_step = staticmethod(lambda fn, warmup=2: GraphedStep(fn, warmup))
    return cls
