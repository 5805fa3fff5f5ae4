"""e2e_slam_b200 -- B200 (sm_100a) kernels for the differentiable-geometry hot path of
End-To-End-Self-Supervised-SLAM, behind the reference's own Python call surface.

Host-side mirror of the reference interface (same names, arguments and error behaviour):
    view_synthesis.BackprojectDepth / Project3D        <- depth_estimation/view_synthesis.py
    losses.SSIM / photometric_loss / ...               <- loss/losses.py
    slam.PointFusion / RGBDImages / Pointclouds / ...  <- gradslam as used by slam/custom_slam.py
    ops.warp_photometric / warp_photometric_loss       <- the fused tier
The compute is in ../csrc (CUDA, C ABI in include/e2e_slam_b200.h); this package holds no fallback.
"""
from . import _lib  # noqa: F401
from . import losses, ops, slam, view_synthesis  # noqa: F401
from .ops import WarpPhotoPlan, photometric_map, ssim_map, warp_photometric, warp_photometric_loss, warp_photometric_multi  # noqa: F401
from .slam import PointFusion, Pointclouds, RGBDImages, image_recover_slam, transform_pointcloud  # noqa: F401

__all__ = ["warp_photometric", "warp_photometric_loss", "warp_photometric_multi", "ssim_map", "photometric_map", "WarpPhotoPlan",
           "PointFusion", "Pointclouds", "RGBDImages", "image_recover_slam", "transform_pointcloud",
           "losses", "ops", "slam", "view_synthesis"]
