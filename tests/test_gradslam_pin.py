"""Pin against the REAL gradslam / chamferdist: golden vectors written by tools/pin_gradslam.py in an environment where those
packages import.  They cannot be produced in the build container (no gradslam, no chamferdist, no network), so until someone
commits tests/golden_gradslam/*.npz every test here is SKIPPED and the fusion / ICP / kNN oracles stay "parity unpinned"."""
import glob
import os

import numpy as np
import pytest

from conftest import ROOT, same_values

PIN_DIR = os.path.join(ROOT, "tests", "golden_gradslam")
HAVE = sorted(os.path.basename(p) for p in glob.glob(os.path.join(PIN_DIR, "*.npz")))
needs_pin = pytest.mark.skipif(not HAVE, reason="parity unpinned: tests/golden_gradslam is empty (run tools/pin_gradslam.py where "
                                                 "gradslam and chamferdist are importable)")


def test_pin_script_reports_missing_packages_cleanly():
    """The generator must say plainly that it cannot pin (exit code 2, nothing written) rather than fall back to a restatement."""
    import subprocess
    import sys
    try:
        import gradslam  # noqa: F401
        pytest.skip("gradslam is importable here: run tools/pin_gradslam.py and commit its output")
    except ImportError:
        pass
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "pin_gradslam.py"), "--out", os.path.join(ROOT, "tests", "_no_such_dir")],
                       capture_output=True, text=True)
    assert r.returncode == 2 and "cannot pin" in r.stdout and not os.path.exists(os.path.join(ROOT, "tests", "_no_such_dir"))


@needs_pin
def test_numpy_oracle_matches_gradslam_pointfusion():
    from oracle import fusion_oracle as fo
    g = np.load(os.path.join(PIN_DIR, "pointfusion_room_6x60x80.npz"))
    o = fo.PointFusionOracle(0.05, 20, 0.6)
    L, H, W = g["depth"].shape
    for s in range(L):
        if s:
            assert np.array_equal(fo.find_active_map_points(o.points, g["K"], g["poses"][s], H, W), g[f"active_rows_{s}"])
        r = o.step(g["depth"][s], g["rgb"][s], g["K"], g["poses"][s])
        assert same_values(r["maps"]["vertex_g"], g[f"vertex_g_{s}"]) == 0 and np.array_equal(r["maps"]["valid"], g[f"valid_{s}"])
        assert o.points.shape == g[f"points_{s}"].shape, f"map size differs at frame {s}"
        for name, ours in (("points", o.points), ("normals", o.normals), ("colors", o.colors), ("ccount", o.ccount)):
            assert np.allclose(ours, g[f"{name}_{s}"], rtol=1e-5, atol=1e-7), (s, name)


@needs_pin
def test_numpy_oracle_matches_chamferdist_ties_and_transform():
    from oracle import fusion_oracle as fo
    g = np.load(os.path.join(PIN_DIR, "knn_ties.npz"))
    d, i = fo.knn1(g["query"], g["ref"])
    assert np.array_equal(i, g["idx"]), "tie rule differs from chamferdist"
    assert np.allclose(d, g["dists"], rtol=1e-6, atol=0)
    t = np.load(os.path.join(PIN_DIR, "transform_pointcloud.npz"))
    assert np.allclose(fo.transform_pointcloud(t["points"], t["T"]), t["out"], rtol=1e-6, atol=1e-6)


@needs_pin
def test_icp_oracle_matches_gradslam():
    from oracle import icp_oracle
    g = np.load(os.path.join(PIN_DIR, "icp_room_pair.npz"))
    T, _ = icp_oracle.point_to_plane_icp(g["src"], g["tgt"], g["nrm"], np.eye(4), 20, 1e-8)
    assert np.allclose(T, g["T_icp"], atol=2e-5)
    Tg, _ = icp_oracle.point_to_plane_gradicp(g["src"], g["tgt"], g["nrm"], np.eye(4), 20, 1e-8, None, 2.0, 1.0, 1.0, 200.0)
    assert np.allclose(Tg, g["T_gradicp"], atol=2e-5)


@needs_pin
@pytest.mark.gpu
def test_cuda_pointfusion_matches_gradslam():
    import torch
    from e2e_slam_b200.slam import PointFusion, Pointclouds, RGBDImages, find_active_map_points
    g = np.load(os.path.join(PIN_DIR, "pointfusion_room_6x60x80.npz"))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    frames = RGBDImages(t(g["rgb"])[None], t(g["depth"])[None, ..., None], t(g["K"]).view(1, 1, 4, 4), t(g["poses"])[None])
    slam, pc = PointFusion(odom="gt", dist_th=0.05, angle_th=20, sigma=0.6, device="cuda"), Pointclouds(device="cuda")
    with torch.no_grad():
        for s in range(g["depth"].shape[0]):
            if s:
                assert np.array_equal(find_active_map_points(pc, frames[:, s]).cpu().numpy(), g[f"active_rows_{s}"])
            pc, _ = slam.step(pc, frames[:, s])
            assert np.allclose(pc.points_list[0].cpu().numpy(), g[f"points_{s}"], rtol=1e-5, atol=1e-7), s
            assert np.allclose(pc.features_list[0][:, 0].cpu().numpy(), g[f"ccount_{s}"], rtol=1e-5, atol=1e-7), s
