"""CPU tests of the C-ABI boundary: the library builds/loads, exports every symbol that
include/vits_mas.h declares, and validates arguments before touching CUDA (no compute here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "vits_mas.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mas_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from vits_b200 import _lib
    declared = _declared_functions()
    assert "mas_maximum_path" in declared and "mas_neg_cent" in declared and "mas_maximum_path_c_host" in declared
    assert sorted(_lib.EXPORTS) == declared


def test_library_exports_every_declared_symbol():
    from vits_b200 import _lib
    L = _lib.lib()
    for name in _declared_functions():
        assert hasattr(L, name), f"libvits_mas.so does not export {name}"
    assert L.mas_abi_version() == 2
    assert L.mas_scratch_status_offset() == 0


def test_scratch_query_and_error_strings():
    from vits_b200 import _lib
    L = _lib.lib()
    n = L.mas_maximum_path_scratch_bytes(64, 1024, 192)
    # >= direction bits (1 bit per cell) + per-frame index + lengths
    assert n >= 64 * 1024 * 192 // 8 + 64 * 1024 * 4 + 64 * 8
    assert n < 8 * 1024 * 1024
    assert L.mas_maximum_path_scratch_bytes(0, 10, 10) == 0
    assert b"scratch" in L.mas_error_string(-4)
    assert L.mas_error_string(0) == b"ok"


def test_argument_validation_without_gpu():
    """Bad arguments are rejected before any CUDA call, so this runs on a CPU-only machine."""
    from vits_b200 import _lib
    L = _lib.lib()
    fake = ctypes.c_void_p(0x1000)   # never dereferenced: validation fails first
    big = 1 << 30
    def call(nc=fake, tys=fake, txs=fake, mask=None, path=fake, idx=None, scratch=fake, sbytes=big,
             B=2, T_y=16, T_x=8, pdt=_lib.MAS_F32):
        return L.mas_maximum_path(nc, tys, txs, mask, 0, 0, 0, 0, path, pdt, idx, scratch, sbytes, B, T_y, T_x, None)
    assert call(B=0) == -1
    assert call(T_x=4096) == -1
    assert call(nc=None) == -3
    assert call(tys=None) == -3            # only one of the two length arrays
    assert call(tys=None, txs=None) == -3  # neither lengths nor mask
    assert call(path=None) == -3           # neither path nor index requested
    assert call(pdt=99) == -2
    assert call(sbytes=16) == -4
    assert call(scratch=ctypes.c_void_p(0x1004)) == -5
    assert L.mas_neg_cent(None, fake, fake, fake, fake, big, 1, 192, 8, 8, None) == -3
    assert L.mas_neg_cent(fake, fake, fake, fake, fake, big, 0, 192, 8, 8, None) == -1
    assert L.mas_maximum_path_c_host(None, fake, fake, fake, 1, 4, 4) == -3
    assert L.mas_maximum_path_host(fake, 99, 0, fake, fake, fake, 1, 4, 4) == -2
    # the streamed stats -> path entry: NULL lengths / misaligned scratch / short scratch are refused up front
    aligned = ctypes.c_void_p(0x10000)
    def fused(tys=fake, scratch=aligned, sbytes=big, B=2):
        return L.mas_stats_to_path(fake, fake, fake, tys, fake, fake, _lib.MAS_F32, None, scratch, sbytes, B, 192, 64, 32, None)
    assert fused(tys=None) == -3
    assert fused(B=0) == -1
    assert fused(scratch=ctypes.c_void_p(0x10010)) == -5
    assert fused(sbytes=1024) == -4
    assert L.mas_stats_to_path_scratch_bytes(64, 192, 1024, 192) > 64 * 1024 * 192 * 4   # holds the neg_cent ring


def test_python_wrapper_rejects_bad_inputs():
    import torch
    from vits_b200.monotonic_align import maximum_path
    nc = torch.zeros(2, 8, 4)
    with pytest.raises(ValueError):
        maximum_path(nc.transpose(1, 2), torch.ones(2, 4, 8))     # not C-contiguous (reference: ValueError)
    with pytest.raises(ValueError):
        maximum_path(torch.zeros(8, 4), torch.ones(8, 4))          # not [b, t_t, t_s]
    with pytest.raises(ValueError):
        maximum_path(torch.zeros(2, 8, 4, dtype=torch.int32), torch.ones(2, 8, 4))


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from vits_b200 import _lib
    from vits_b200.monotonic_align import maximum_path
    with pytest.raises(_lib.MasError):
        maximum_path(torch.zeros(1, 8, 4), torch.ones(1, 8, 4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "vits_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "mas_oracle" not in text, f


def test_lengths_from_a_cpu_mask_match_the_reference_expression():
    """The host side of `maximum_path(neg_cent_cpu, mask_cpu)`: `mask.sum(1)[:, 0]` / `mask.sum(2)[:, 0]`
    (monotonic_align/__init__.py:17-18) from column 0 / row 0 only, for every mask dtype a caller can hand in -- exact
    counts also where a half-precision sum would round (t_y > 2048)."""
    import numpy as np
    import torch
    from vits_b200.monotonic_align import _lengths_from_mask_cpu
    T_y, T_x = 4100, 40
    ty = torch.tensor([4099, 2049, 17, 0])
    tx = torch.tensor([40, 33, 1, 0])
    prefix = (torch.arange(T_y)[None, :] < ty[:, None])[:, :, None] & (torch.arange(T_x)[None, :] < tx[:, None])[:, None, :]
    for dt in (torch.float32, torch.float64, torch.float16, torch.bfloat16, torch.bool, torch.uint8, torch.int64):
        m = prefix.to(dt)
        got_y, got_x = _lengths_from_mask_cpu(m)
        assert got_y.dtype == torch.int32 and got_x.dtype == torch.int32
        assert got_y.tolist() == ty.tolist() and got_x.tolist() == tx.tolist(), dt
    # and literally the reference's expression where that is exact
    m = prefix.float()
    ref_y = m.sum(1)[:, 0].numpy().astype(np.int32)
    ref_x = m.sum(2)[:, 0].numpy().astype(np.int32)
    got_y, got_x = _lengths_from_mask_cpu(m)
    assert np.array_equal(got_y.numpy(), ref_y) and np.array_equal(got_x.numpy(), ref_x)
    # a strided (non-contiguous) mask view is accepted as it is
    wide = torch.zeros(4, T_y, 2 * T_x)
    wide[:, :, ::2] = m
    got_y, got_x = _lengths_from_mask_cpu(wide[:, :, ::2])
    assert got_y.tolist() == ty.tolist() and got_x.tolist() == tx.tolist()
