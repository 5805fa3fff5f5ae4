"""CPU tests of the drop-in wiring: after patch.install() the reference's own import statements resolve to
this package (no gradslam / chamferdist / matplotlib needed)."""
import importlib
import sys


def test_install_routes_reference_imports():
    import e2e_slam_b200
    from e2e_slam_b200 import patch
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("gradslam", "chamferdist", "loss", "slam", "depth_estimation")}
    try:
        patch.install()
        ns = {}
        exec("from gradslam.slam import PointFusion, ICPSLAM\n"
             "from gradslam import Pointclouds, RGBDImages\n"
             "from gradslam.slam.fusionutils import find_active_map_points\n"
             "from gradslam.geometry.geometryutils import transform_pointcloud\n"
             "from chamferdist import ChamferDistance\n"
             "from chamferdist.chamfer import knn_points\n", ns)
        assert ns["PointFusion"] is e2e_slam_b200.slam.PointFusion
        assert ns["RGBDImages"] is e2e_slam_b200.slam.RGBDImages
        assert ns["knn_points"] is e2e_slam_b200.losses.knn_points
        assert ns["transform_pointcloud"] is e2e_slam_b200.slam.transform_pointcloud
        losses = sys.modules["loss.losses"]
        for name in ("SSIM", "knn_points_loss", "color_points_loss", "geometric_consistency_loss", "photometric_loss",
                     "disparity_smoothness_loss", "depth_reguralizer", "depth_gt_loss", "depth_metrics", "compute_depth_errors"):
            assert hasattr(losses, name), name                      # every public name of the reference's loss/losses.py
        vs = sys.modules["depth_estimation.view_synthesis"]
        assert vs.BackprojectDepth is e2e_slam_b200.view_synthesis.BackprojectDepth and hasattr(vs, "Project3D")
        assert sys.modules["slam.custom_slam"].image_recover_slam is e2e_slam_b200.slam.image_recover_slam
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in ("gradslam", "chamferdist", "loss", "slam", "depth_estimation")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_reference_signatures_are_mirrored():
    """Same constructor / call signatures as the reference."""
    import inspect
    from e2e_slam_b200 import losses, slam, view_synthesis
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(view_synthesis.BackprojectDepth.__init__) == ["self", "batch_size", "height", "width"]
    assert sig(view_synthesis.BackprojectDepth.forward) == ["self", "depth", "inv_K"]
    assert sig(view_synthesis.Project3D.__init__) == ["self", "batch_size", "height", "width", "eps"]
    assert sig(view_synthesis.Project3D.forward) == ["self", "points", "K", "T", "geometric"]
    assert sig(losses.SSIM.forward) == ["self", "x", "y"]
    assert sig(losses.photometric_loss) == ["ssim", "prediction", "target"]
    assert sig(losses.knn_points_loss) == ["gt_pointcloud", "noisy_pointcloud"]
    assert sig(losses.color_points_loss) == ["gt_pointcloud_color", "noisy_pointcloud_color", "indexes"]
    assert sig(losses.geometric_consistency_loss) == ["outputs", "frame", "device"]
    assert sig(losses.disparity_smoothness_loss) == ["disp", "img"]
    assert sig(losses.depth_reguralizer) == ["initial_depth", "refined_depth", "loss_func"]
    assert sig(losses.depth_gt_loss) == ["prediction", "sparse_groundtruth", "sparse_mask"]
    assert sig(losses.depth_metrics) == ["dataset", "gt", "pred"]
    assert sig(slam.image_recover_slam) == ["noisy_rgbd", "slam", "device"]
    assert sig(slam.PointFusion.step) == ["self", "pointclouds", "live_frame", "prev_frame", "inplace"]
    assert sig(slam.PointFusion.__init__)[:5] == ["self", "odom", "dist_th", "angle_th", "sigma"]


def test_rgbdimages_container_cpu():
    import pytest
    import torch
    from e2e_slam_b200.slam import RGBDImages
    r = RGBDImages(torch.rand(2, 3, 4, 5, 3), torch.rand(2, 3, 4, 5, 1), torch.eye(4).repeat(2, 1, 1, 1))
    assert r.shape == (2, 3, 4, 5) and r.poses is None
    f = r[:, 1]
    assert f.shape == (2, 1, 4, 5) and torch.equal(f.rgb_image[:, 0], r.rgb_image[:, 1])
    f.poses = torch.eye(4).view(1, 1, 4, 4).repeat(2, 1, 1, 1)
    assert f.detach().poses.shape == (2, 1, 4, 4)
    with pytest.raises(ValueError):
        f.poses = torch.eye(4)
    with pytest.raises(TypeError):
        RGBDImages([1], torch.rand(1, 1, 2, 2, 1), torch.eye(4).view(1, 1, 4, 4))
