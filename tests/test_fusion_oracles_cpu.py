"""CPU cross-check of the two independently written PointFusion restatements (oracle/fusion_oracle.py: numpy index arrays +
lexsort; oracle/fusion_oracle_torch.py: torch row tensors + torch.unique(dim=0), the way gradslam itself is written).  Neither is
pinned to gradslam (not installable here; tools/pin_gradslam.py regenerates goldens from it when it is), but two restatements
written against the same frozen semantics must agree bit for bit -- rows, index maps, append order, merged map."""
import numpy as np
import pytest
import torch

from conftest import same_values


@pytest.mark.parametrize("L,H,W,holes", [(6, 48, 64, 0.0), (5, 60, 80, 0.15)])
def test_numpy_and_torch_restatements_agree(L, H, W, holes):
    from oracle import fusion_oracle as fo
    from oracle import fusion_oracle_torch as ft
    depth, rgb, K, poses = fo.synthetic_room_sequence(L, H, W, seed=3)
    if holes:
        depth = depth * (np.random.default_rng(4).random(depth.shape) >= holes)
    depth, rgb = depth.astype(np.float32), rgb.astype(np.float32)
    a, b = fo.PointFusionOracle(0.05, 20, 0.6), ft.PointFusionOracleTorch(0.05, 20, 0.6)
    matched = 0
    for s in range(L):
        n_before = len(a.points)
        active_np = fo.find_active_map_points(a.points, K, poses[s], H, W)
        oa = a.step(depth[s], rgb[s], K, poses[s])
        ob = b.step(depth[s], rgb[s], K, poses[s])
        for k in ("vertex_g", "normal_g", "alpha"):
            assert same_values(oa["maps"][k], ob["maps"][k].numpy()) == 0, (s, k)
        assert np.array_equal(active_np, ob["active"].numpy()), f"active rows differ at frame {s}"
        assert np.array_equal(oa["rows"], ob["rows"].numpy()), f"correspondence rows differ at frame {s}"
        assert np.array_equal(oa["append_slot"] >= 0, ob["appended"].numpy()), f"appended pixels differ at frame {s}"
        assert same_values(a.points, b.points.numpy()) == 0 and same_values(a.normals, b.normals.numpy()) == 0
        assert same_values(a.colors, b.colors.numpy()) == 0 and same_values(a.ccount, b.ccount.numpy()) == 0
        matched += len(oa["rows"])
        assert len(active_np) <= n_before
    assert matched > 0


def test_active_rows_are_ordered_and_in_frame():
    from oracle import fusion_oracle as fo
    depth, rgb, K, poses = fo.synthetic_room_sequence(3, 48, 64, seed=1)
    o = fo.PointFusionOracle()
    o.step(depth[0].astype(np.float32), rgb[0].astype(np.float32), K, poses[0])
    rows = fo.find_active_map_points(o.points, K, poses[2], 48, 64)
    assert rows.dtype == np.int64 and rows.shape[1] == 4 and len(rows) > 0
    assert np.all(np.diff(rows[:, 1]) > 0) and np.all(rows[:, 0] == 0)
    assert rows[:, 2].min() >= 0 and rows[:, 2].max() < 48 and rows[:, 3].min() >= 0 and rows[:, 3].max() < 64


def test_rgbdimages_negative_index():
    """gradslam's container accepts negative indices; frames[:, -1] used to come back as an EMPTY sequence."""
    from e2e_slam_b200.slam import RGBDImages
    B, L, H, W = 2, 3, 4, 5
    rgb = torch.arange(B * L * H * W * 3, dtype=torch.float32).view(B, L, H, W, 3)
    depth = torch.ones(B, L, H, W, 1)
    K = torch.eye(4).view(1, 1, 4, 4).repeat(B, 1, 1, 1)
    poses = torch.eye(4).view(1, 1, 4, 4).repeat(B, L, 1, 1)
    f = RGBDImages(rgb, depth, K, poses)
    last = f[:, -1]
    assert last.shape == (B, 1, H, W) and torch.equal(last.rgb_image, rgb[:, 2:3]) and torch.equal(last.poses, poses[:, 2:3])
    assert f[-1].shape == (1, L, H, W) and torch.equal(f[-1].rgb_image, rgb[1:2])
    assert f[-2, -3].shape == (1, 1, H, W) and torch.equal(f[-2, -3].rgb_image, rgb[0:1, 0:1])
    with pytest.raises(IndexError):
        f[:, 3]
    with pytest.raises(IndexError):
        f[-3]
