"""ctypes binding of libe2e_slam_b200.so (the C ABI declared in include/e2e_slam_b200.h).

There is deliberately no fallback of any kind: if the shared library is missing, cannot be loaded,
or a tensor is not a CUDA fp32 tensor, the call raises.  PyTorch is used for device memory, streams
and autograd plumbing only.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "..", "lib", "libe2e_slam_b200.so")

c_f32p = ctypes.c_void_p      # device pointers travel as integers
c_strides = ctypes.POINTER(ctypes.c_int64)
_I, _F, _P, _S, _LL, _SZ = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, c_strides, ctypes.c_longlong, ctypes.c_size_t

_SIGNATURES = {
    "e2e_abi_version": (ctypes.c_int, []),
    "e2e_last_error": (ctypes.c_char_p, []),
    "e2e_launch_count": (ctypes.c_ulonglong, []),
    "e2e_prepare_divisor": (_I, [_F, _P]),
    "e2e_warp_photo_workspace_bytes": (_SZ, [_I, _I, _I]),
    "e2e_warp_photo_fwd": (_I, [_P, _P, _P, _P, _P, _S, _P, _S, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "e2e_warp_photo_bwd": (_I, [_P, _P, _P, _P, _P, _S, _P, _S, _I, _I, _I, _I, _I, _F, _P, _P, _F, _P, _P, _S, _P, _P, _SZ, _P]),
    "e2e_warp_photo_vg_workspace_bytes": (_SZ, [_I, _I, _I]),
    "e2e_warp_photo_vg": (_I, [_P, _P, _P, _P, _P, _S, _P, _S, _I, _I, _I, _I, _I, _F, _P, _P, _P, _S, _P, _P, _SZ, _P]),
    "e2e_warp_photo_vg_multi_workspace_bytes": (_SZ, [_I, _I, _I, _I]),
    "e2e_warp_photo_vg_multi": (_I, [_P, _P, _P, _P, _P, _S, _P, _S, _I, _I, _I, _I, _I, _I, _F, _P, _P, _P, _S, _P, _P, _SZ, _P]),
    "e2e_warp_photo_vg_disp": (_I, [_P, _P, _P, _P, _P, _P, _S, _P, _S, _I, _I, _I, _I, _I, _F, _P, _P, _P, _S, _P, _P, _SZ, _P]),
    "e2e_scale_by_scalar": (_I, [_P, _LL, _P, _LL, _P, _LL, _P, _P]),
    "e2e_u8_to_unit": (_I, [_P, _LL, _P, _P]),
    "e2e_disp_to_depth_fwd": (_I, [_P, _P, _LL, _P, _P]),
    "e2e_disp_to_depth_bwd": (_I, [_P, _P, _P, _LL, _P, _P]),
    "e2e_select_workspace_bytes": (_SZ, []),
    "e2e_select_kth": (_I, [_P, _LL, _LL, _P, _P, _SZ, _P]),
    "e2e_dual_disparity_fwd": (_I, [_P, _P, _P, _I, _I, _P, _P]),
    "e2e_dual_disparity_bwd": (_I, [_P, _P, _I, _I, _P, _P, _P]),
    "e2e_warp_photo_bwd_cond": (_I, [_P, _P, _P, _P, _P, _S, _P, _S, _I, _I, _I, _I, _I, _F, _P, _P, _F, _P, _P, _P, _S, _P, _P, _SZ, _P]),
    "e2e_warp_photo_vg_map": (_I, [_P, _P, _P, _P, _P, _S, _P, _S, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P, _P, _S, _P, _P, _SZ, _P]),
    "e2e_upstream_uniform": (_I, [_P, _LL, ctypes.c_double, _P, _P]),
    "e2e_scale_or_zero": (_I, [_P, _LL, _P, _LL, _P, _LL, _P, _P]),
    "e2e_ssim_fwd": (_I, [_P, _S, _P, _S, _I, _I, _I, _I, _P, _P, _P]),
    "e2e_ssim_bwd": (_I, [_P, _S, _P, _S, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "e2e_backproject_fwd": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "e2e_backproject_bwd": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "e2e_project3d_fwd": (_I, [_P, _P, _P, _I, _I, _I, _F, _P, _P, _P, _P]),
    "e2e_project3d_bwd": (_I, [_P, _P, _P, _I, _I, _I, _F, _P, _P, _P, _P, _P, _SZ, _P]),
    "e2e_grid_sample_fwd": (_I, [_P, _S, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "e2e_grid_sample_bwd": (_I, [_P, _P, _S, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _S, _P, _P]),
    "e2e_reduce_workspace_bytes": (_SZ, [_LL]),
    "e2e_smooth_fwd": (_I, [_P, _P, _S, _I, _I, _I, _P, _P, _SZ, _P]),
    "e2e_smooth_bwd": (_I, [_P, _P, _S, _I, _I, _I, _P, _P, _P, _SZ, _P]),
    "e2e_smooth_vg_workspace_bytes": (_SZ, [_I, _I, _I]),
    "e2e_smooth_vg": (_I, [_P, _P, _S, _I, _I, _I, _P, _P, _P, _P, _SZ, _P]),
    "e2e_smooth_apply": (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "e2e_smooth_vg_raw": (_I, [_P, _P, _S, _I, _I, _I, _P, _P, _P, _P, _SZ, _P]),
    "e2e_sparse_l1_fwd": (_I, [_P, _P, _P, _LL, _P, _P, _SZ, _P]),
    "e2e_sparse_l1_bwd": (_I, [_P, _P, _P, _LL, _P, _P, _P]),
    "e2e_depth_reg_fwd": (_I, [_P, _P, _LL, _I, _P, _P, _SZ, _P]),
    "e2e_depth_reg_bwd": (_I, [_P, _P, _LL, _I, _P, _P, _P]),
    "e2e_geometric_fwd": (_I, [_P, _P, _P, _LL, _P, _P, _SZ, _P]),
    "e2e_geometric_bwd": (_I, [_P, _P, _P, _LL, _P, _P, _P, _P, _P]),
    "e2e_min_composite_fwd": (_I, [_P, _I, _LL, _P, _P, _P, _SZ, _P]),
    "e2e_min_composite_bwd": (_I, [_P, _I, _LL, _P, _P, _P]),
    "e2e_rgbd_maps": (_I, [_P, _P, _P, _P, _I, _I, _F, _P, _P, _P, _P, _P]),
    "e2e_rgbd_maps_bwd": (_I, [_P, _P, _P, _I, _I, _F, _P, _P, _P, _P, _P]),
    "e2e_fusion_associate": (_I, [_P, _P, _P, _P, _LL, _P, _P, _P, _P, _I, _I, _F, _F, _P, _P, _P, _P]),
    "e2e_fusion_active_points_workspace_bytes": (_SZ, [_LL]),
    "e2e_fusion_active_points": (_I, [_P, _LL, _P, _P, _I, _I, _LL, _P, _P, _P, _SZ, _P]),
    "e2e_fusion_workspace_bytes": (_SZ, [_I, _I]),
    "e2e_fusion_merge_append": (_I, [_P, _P, _P, _P, _P, _LL, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _SZ, _P]),
    "e2e_knn1_grid_workspace_bytes": (_SZ, [_LL]),
    "e2e_knn1_grid_fwd": (_I, [_P, _P, _P, _LL, _LL, _P, _P, _P, _SZ, _P]),
    "e2e_knn1_grid_build": (_I, [_P, _LL, _P, _SZ, _P]),
    "e2e_knn1_grid_query": (_I, [_P, _P, _LL, _LL, _P, _P, _P, _P]),
    "e2e_icp_workspace_bytes": (_SZ, [_LL, _LL]),
    "e2e_icp_point_to_plane": (_I, [_P, _LL, _P, _P, _LL, _P, _I, _F, _F, _I, _F, _F, _F, _F, _P, _P, _P, _P, _SZ, _P]),
    "e2e_icp_history_bytes": (_SZ, [_LL, _I]),
    "e2e_icp_point_to_plane_saved": (_I, [_P, _LL, _P, _P, _LL, _P, _I, _F, _F, _I, _F, _F, _F, _F, _P, _P, _P, _SZ, _P, _SZ, _P]),
    "e2e_icp_backward_workspace_bytes": (_SZ, [_LL]),
    "e2e_icp_backward": (_I, [_P, _LL, _P, _P, _LL, _P, _I, _F, _F, _I, _F, _F, _F, _F, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "e2e_fusion_sequence_workspace_bytes": (_SZ, [_I, _I, _LL]),
    "e2e_fusion_sequence": (_I, [_P, _P, _P, _P, _I, _I, _I, _F, _F, _F, _P, _P, _P, _P, _P, _LL, _LL, _P, _SZ, _P]),
    "e2e_fusion_sequence_batch_workspace_bytes": (_SZ, [_I, _I, _I, _LL]),
    "e2e_fusion_sequence_batch": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _F, _P, _P, _P, _P, _P, _LL, _P, _SZ, _P]),
    "e2e_fusion_merge_append_bwd": (_I, [_P] * 11 + [_I, _I] + [_P] * 7),
    "e2e_multimem_allreduce_avg": (_I, [_P, _LL, _I, _I, _I, _P]),
    "e2e_knn1_fwd": (_I, [_P, _P, _P, _LL, _LL, _P, _P, _P]),
    "e2e_knn1_bwd": (_I, [_P, _P, _P, _LL, _LL, _P, _P, _P, _P, _P]),
    "e2e_transform_points_fwd": (_I, [_P, _P, _LL, _P, _P]),
    "e2e_transform_points_bwd": (_I, [_P, _P, _LL, _P, _P]),
    "e2e_color_points_workspace_bytes": (_SZ, [_LL]),
    "e2e_color_points_fwd": (_I, [_P, _P, _P, _LL, _P, _P, _SZ, _P]),
    "e2e_color_points_bwd": (_I, [_P, _P, _P, _LL, _P, _P, _P, _P]),
}

_lib = None


class E2ELibraryError(RuntimeError):
    pass


def exported_symbols():
    """Every symbol include/e2e_slam_b200.h declares (used by the CPU-side ABI test)."""
    return sorted(_SIGNATURES)


def lib():
    global _lib
    if _lib is None:
        path = os.path.abspath(os.environ.get("E2E_LIB_PATH") or LIB_PATH)      # override: kernel-tuning experiments only
        if not os.path.exists(path):
            raise E2ELibraryError(
                f"{path} not found: build it with `python end-to-end-self-supervised-slam_b200/build.py` "
                "(or __graft_entry__.build()).  There is no CPU or PyTorch fallback.")
        handle = ctypes.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)     # AttributeError here = header and library out of sync
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().e2e_last_error().decode() or "no message"
        raise E2ELibraryError(f"{what} failed (code {rc}): {msg}")


def launch_count():
    return int(lib().e2e_launch_count())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a CUDA fp32/int64/uint8 tensor (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    if not t.is_cuda:
        raise E2ELibraryError("e2e_slam_b200 kernels need CUDA tensors (there is no CPU path)")
    return ctypes.c_void_p(t.data_ptr())


def f32(t, name="tensor"):
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (the reference path is fp32 throughout), got {t.dtype}")
    if not t.is_cuda:
        raise E2ELibraryError(f"{name} must be a CUDA tensor (there is no CPU path)")
    return t


def strides4(t):
    if t.dim() != 4:
        raise ValueError(f"expected a 4-D (B,C,H,W) tensor, got shape {tuple(t.shape)}")
    return (ctypes.c_int64 * 4)(*t.stride())


_prepared = set()


def prepare_divisors(*ds):
    """Verify constant divisors once per process, outside any CUDA-graph capture (it synchronises)."""
    for d in ds:
        d = float(d)
        if d in _prepared:
            continue
        rc = lib().e2e_prepare_divisor(ctypes.c_float(d), stream_ptr())
        if rc < 0:
            check(rc, "e2e_prepare_divisor")
        _prepared.add(d)
