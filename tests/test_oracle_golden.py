"""CPU tests: the oracles against the golden vectors produced by the real reference
(tools/make_golden.py).  These pin the oracles; the GPU tests then compare CUDA with them."""
import numpy as np
import pytest
import torch

from conftest import rel_max, same_values
from oracle import analytic_backward, c_oracle, torch_oracle


def _t(g, k):
    return torch.from_numpy(g[k])


def test_c_oracle_bit_exact_vs_reference(golden):
    g = golden
    o = c_oracle.warp_photo_fwd(g["depth"], g["inv_K"], g["K"], g["T"], g["colors"][:, 0], g["colors"][:, 1],
                                str(g["padding_mode"]), bool(g["use_mask"]))
    for k in ("pix", "valid", "syn", "ssim", "loss_map"):
        assert same_values(o[k], g[k]) == 0, k


def test_torch_oracle_bit_exact_vs_reference(golden):
    g = golden
    r = torch_oracle.fwd_bwd(_t(g, "depth"), _t(g, "inv_K"), _t(g, "K"), _t(g, "T"), _t(g, "colors")[:, 0],
                             _t(g, "colors")[:, 1], str(g["padding_mode"]), bool(g["use_mask"]))
    for k in ("pix", "valid", "syn", "loss_map", "g_depth", "g_src", "g_T"):
        assert same_values(r[k].numpy(), g[k]) == 0, k
    assert abs(float(r["loss"]) - float(g["loss"])) <= 1e-7 * abs(float(g["loss"]))


def test_torch_oracle_fp64_matches_reference_fp64(golden):
    g = golden
    r = torch_oracle.fwd_bwd(_t(g, "depth"), _t(g, "inv_K"), _t(g, "K"), _t(g, "T"), _t(g, "colors")[:, 0],
                             _t(g, "colors")[:, 1], str(g["padding_mode"]), bool(g["use_mask"]), dtype=torch.float64)
    for k in ("g_depth", "g_src", "g_T"):
        assert rel_max(r[k].numpy(), g[k + "_f64"]) < 1e-12, k


def test_closed_form_backward_matches_autograd_fp64(golden):
    """The formulas the CUDA backward implements (oracle/analytic_backward.py) vs autograd in float64.
    tum_flat has exactly flat regions where the SSIM clamp sits on its kink (value == 0 up to rounding):
    there the gradient w.r.t. the image is decided by rounding noise, so g_src is skipped for it."""
    g = golden
    r = analytic_backward.backward(g["depth"], g["inv_K"], g["K"], g["T"], g["colors"][:, 0], g["colors"][:, 1],
                                   str(g["padding_mode"]), bool(g["use_mask"]))
    assert abs(r["loss"] - float(g["loss_f64"])) < 1e-12
    assert rel_max(r["g_depth"], g["g_depth_f64"]) < 1e-10
    assert rel_max(r["g_T"], g["g_T_f64"]) < 1e-10
    if not bool((g["colors"] == 1.0).any()):
        assert rel_max(r["g_src"], g["g_src_f64"]) < 1e-10


def test_small_losses_oracle_vs_reference(golden):
    g = golden
    disp = _t(g, "disp").clone().requires_grad_(True)
    tgt = _t(g, "colors")[:, 1].permute(0, 3, 1, 2)
    s = torch_oracle.smoothness(disp, tgt)
    s.backward()
    assert same_values(s.detach().numpy(), g["smooth"]) == 0
    assert same_values(disp.grad.numpy(), g["g_disp_smooth"]) == 0
    if "gt_loss" in g.files:
        pred = _t(g, "pred_depth").clone().requires_grad_(True)
        l = torch_oracle.sparse_gt_l1(pred, _t(g, "sparse_gt"), _t(g, "sparse_mask"))
        l.backward()
        assert same_values(l.detach().numpy(), g["gt_loss"]) == 0
        assert same_values(pred.grad.numpy(), g["g_pred_gt"]) == 0
    for kind in ("l1", "l2"):
        b = (_t(g, "depth") * 1.05 + 0.01).clone().requires_grad_(True)
        l = torch_oracle.depth_regulariser(_t(g, "depth"), b, kind)
        l.backward()
        assert same_values(l.detach().numpy(), g["reg_" + kind]) == 0
        assert same_values(b.grad.numpy(), g["g_reg_" + kind]) == 0
    gl = torch_oracle.geometric_consistency(_t(g, "warped_depth"), _t(g, "interp_depth"), _t(g, "valid"))
    assert same_values(gl.numpy(), g["geo_loss"]) == 0


@pytest.mark.parametrize("mr,am", [(False, False), (True, False), (False, True), (True, True)])
def test_torch_oracle_multi_source_objective_matches_reference(composite_golden, mr, am):
    """oracle/torch_oracle.photometric_objective (mean over frames / min-reprojection / auto-masking, train_depth.py:615-660, 729-750)
    against goldens produced with the reference's own modules (tools/make_golden_composite.py): same winning candidate at every pixel,
    same loss bit for bit, depth gradient to 1e-5 (the reference back-projects once for all source frames,
    the oracle once per frame: autograd accumulates in a different order)."""
    import torch
    from oracle import torch_oracle as to
    g = composite_golden
    tag = f"mr{int(mr)}_am{int(am)}"
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    depth = t(g["depth"]).requires_grad_(True)
    colors = t(g["colors"])
    S = g["T"].shape[1]
    srcs = [colors[:, 1 + s].permute(0, 3, 1, 2) for s in range(S)]
    torch.set_num_threads(1)                                   # the generator fixed torch's reduction tree the same way
    loss, index = to.photometric_objective(depth, t(g["inv_K"]), t(g["K"]), [t(g["T"][:, s]) for s in range(S)], srcs,
                                           colors[:, 0].permute(0, 3, 1, 2), str(g["padding_mode"]), bool(g["use_mask"]), mr, am, t(g["noise"]))
    loss.backward()
    if index is not None:
        assert np.array_equal(index.numpy(), g[f"index_{tag}"])
    assert float(loss) == float(g[f"loss_{tag}"])
    assert rel_max(depth.grad.numpy(), g[f"g_depth_{tag}"]) <= 1e-5
