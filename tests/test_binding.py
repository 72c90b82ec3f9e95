"""The reference-side native binding, compiled for real: vits_b200/binding/core.pyx keeps
the signature of the reference's `maximum_path_c` (monotonic_align/core.pyx:38) and calls libvits_mas.so; it is loaded
beneath the reference's UNMODIFIED wrapper (monotonic_align/__init__.py, copied byte for byte into the git-ignored
baseline/_ref/binding by tools/install_reference.py) and must reproduce the goldens generated from the real reference."""
import filecmp
import glob
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from helpers import index_to_path, sha_path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "baseline", "_ref", "binding", "monotonic_align")


def _load_bound_package():
    if not glob.glob(os.path.join(PKG, "monotonic_align", "core*.so")):
        pytest.skip("baseline/_ref/binding not built (python tools/install_reference.py)")
    name = "ref_monotonic_align_on_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG, "__init__.py"), submodule_search_locations=[PKG])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)      # runs `from .monotonic_align.core import maximum_path_c` (__init__.py:4)
    return mod


def test_binding_is_the_reference_wrapper_over_our_core():
    mod = _load_bound_package()
    ref_init = "/root/reference/monotonic_align/__init__.py"
    if os.path.exists(ref_init):
        assert filecmp.cmp(ref_init, os.path.join(PKG, "__init__.py"), shallow=False), "wrapper must be the reference's, unmodified"
    core = sys.modules[mod.__name__ + ".monotonic_align.core"]
    assert core.__file__.startswith(PKG) and callable(core.maximum_path_c)
    # the compiled core resolves its compute entry from libvits_mas.so, not from any oracle library
    with open(core.__file__, "rb") as f:
        blob = f.read()
    assert b"mas_maximum_path_c_host" in blob and b"libvits_mas.so" in blob and b"mas_oracle" not in blob
    # same typed-memoryview contract as core.pyx:38: non-contiguous input is rejected at the boundary
    with pytest.raises(ValueError):
        core.maximum_path_c(np.zeros((1, 4, 6), np.int32)[:, :, ::2], np.zeros((1, 4, 3), np.float32),
                            np.array([4], np.int32), np.array([3], np.int32))


@pytest.mark.gpu
def test_reference_wrapper_over_binding_matches_goldens(mas_golden):
    mod = _load_bound_package()
    for name, c in mas_golden.items():
        nc = torch.from_numpy(c["neg_cent"].copy())
        B, T_y, T_x = nc.shape
        ym = torch.arange(T_y)[None, :] < torch.as_tensor(c["t_ys"])[:, None]
        xm = torch.arange(T_x)[None, :] < torch.as_tensor(c["t_xs"])[:, None]
        mask = (ym[:, :, None] & xm[:, None, :]).float()
        for dev in ("cpu", "cuda"):       # the reference's wrapper moves CUDA tensors through the host itself (:14,:20)
            out = mod.maximum_path(nc.to(dev), mask.to(dev))
            assert out.dtype == torch.float32 and out.device.type == dev
            got = out.cpu().numpy().astype(np.int8)
            np.testing.assert_array_equal(got, index_to_path(c["index"].astype(np.int32), T_x), err_msg=name)
            assert sha_path(got) == c["sha256"], name
        assert torch.equal(nc, torch.from_numpy(c["neg_cent"])), "input must not be modified"
