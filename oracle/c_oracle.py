"""ctypes loader (and gcc build recipe) for oracle/warp_photo_oracle.c -- TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "warp_photo_oracle.c")
_LIB = os.path.join(_HERE, "libwarp_photo_oracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-mfma", "-ffp-contract=off", "-fno-fast-math",
               _SRC, "-o", _LIB, "-lm"]
        subprocess.check_call(cmd)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.e2e_oracle_warp_photo_fwd.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def warp_photo_fwd(depth, inv_K, K, T, src_cl, tgt_cl, padding_mode="border", use_mask=True, eps=1e-7):
    """numpy float32 in (depth (B,1,H,W), 4x4s (B,4,4), channels-last images (B,H,W,3)) -> dict of numpy."""
    f = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    depth, inv_K, K, T, src_cl, tgt_cl = map(f, (depth, inv_K, K, T, src_cl, tgt_cl))
    B, _, H, W = depth.shape
    out = dict(pix=np.empty((B, H, W, 2), np.float32), valid=np.empty((B, 1, H, W), np.float32),
               syn=np.empty((B, 3, H, W), np.float32), ssim=np.empty((B, 3, H, W), np.float32),
               loss_map=np.empty((B, 1, H, W), np.float32))
    rc = lib().e2e_oracle_warp_photo_fwd(_p(depth), _p(inv_K), _p(K), _p(T), _p(src_cl), _p(tgt_cl),
                                          B, H, W, int(padding_mode == "border"), int(bool(use_mask)),
                                          ctypes.c_float(eps), _p(out["pix"]), _p(out["valid"]), _p(out["syn"]),
                                          _p(out["ssim"]), _p(out["loss_map"]))
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    return out
