"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol
that include/e2e_slam_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "e2e_slam_b200.h")


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__
    __graft_entry__.build()
    from e2e_slam_b200 import _lib
    return _lib


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(e2e_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_symbols():
    syms = declared_symbols()
    assert "e2e_warp_photo_fwd" in syms and "e2e_warp_photo_bwd" in syms and len(syms) >= 9


def test_library_exports_every_declared_symbol(built_lib):
    handle = ctypes.CDLL(os.path.abspath(built_lib.LIB_PATH))
    missing = [s for s in declared_symbols() if not hasattr(handle, s)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_host_binding_covers_the_header(built_lib):
    assert sorted(built_lib.exported_symbols()) == declared_symbols()


def test_abi_version_and_launch_counter(built_lib):
    lib = built_lib.lib()
    assert lib.e2e_abi_version() == 1
    assert built_lib.launch_count() == 0          # nothing launched on a CPU-only box


def test_cpu_tensors_are_rejected_not_emulated(built_lib):
    import torch
    import e2e_slam_b200 as e2e
    from e2e_slam_b200.synthetic import make_pairs
    d = make_pairs(1, 8, 8)
    src, tgt = d["colors"][:, 0].permute(0, 3, 1, 2), d["colors"][:, 1].permute(0, 3, 1, 2)
    with pytest.raises(RuntimeError):
        e2e.warp_photometric(d["depth"], d["inv_K"], d["K"], d["T"], src, tgt)
