"""Drop-in for the reference package ``monotonic_align`` (monotonic_align/__init__.py:7-20).

``maximum_path(neg_cent, mask)`` keeps the reference's name, positional arguments and result
contract -- a new tensor ``[B, T_y, T_x]`` with ``neg_cent``'s dtype on ``neg_cent``'s device,
values exactly 0/1, inputs not modified -- but runs the dynamic program, the backtrack and the
path materialisation as sm_100a CUDA kernels (vits_b200/csrc/mas_path.cu) behind the C ABI in
include/vits_mas.h.  On CUDA tensors it launches on the caller's current stream and never
synchronises (the reference does a blocking D2H, a serial CPU loop and an H2D).

``SynthesizerTrn`` binds the name at import (SynthesizerTrn.py:16), so either put the directory
that contains this package ahead of the reference's on ``sys.path`` or assign
``SynthesizerTrn.maximum_path = vits_b200.monotonic_align.maximum_path``.

There is no CPU implementation here: CPU tensors are shipped through the host-buffer C entry
(``mas_maximum_path_c_host``, the twin of core.pyx:38), which still needs a CUDA device.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .. import _lib

__all__ = ["maximum_path", "maximum_path_index", "maximum_path_from_lengths", "last_status", "status_nosync",
           "StatusError"]

_DT = {
    torch.float32: _lib.MAS_F32, torch.float16: _lib.MAS_F16, torch.bfloat16: _lib.MAS_BF16,
    torch.float64: _lib.MAS_F64, torch.uint8: _lib.MAS_U8, torch.bool: _lib.MAS_U8, torch.int8: _lib.MAS_I8,
    torch.int16: _lib.MAS_I16, torch.int32: _lib.MAS_I32, torch.int64: _lib.MAS_I64,
}

_scratch: Dict[Tuple[int, int], torch.Tensor] = {}
_CHECK = os.environ.get("VITS_MAS_CHECK", "0") == "1"


class StatusError(ValueError):
    """Raised (only when checking is requested) for lengths the reference leaves undefined."""


def _scratch_for(device: torch.device, stream: int, nbytes: int) -> torch.Tensor:
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        # zero-initialised: the first word is the sticky status word (carried over when the buffer grows)
        new = torch.zeros(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        if buf is not None:
            new[:4].copy_(buf[:4])
        _scratch[key] = buf = new
    return buf


def _mirror(device: torch.device):
    """The library's host-mapped status words of `device` (ctypes int32[4]) or None."""
    with torch.cuda.device(device):
        p = _lib.lib().mas_status_mirror()
    return p if p else None


def status_nosync(device=None, reset: bool = False) -> int:
    """MAS_STATUS_* bits raised by any call on ``device`` so far, read from the library's host-mapped mirror:
    no synchronisation, no copy (bits of kernels still in flight show up once they have run).  The dense status
    word read by :func:`last_status` is per stream and needs a sync; this one is what a training loop can poll
    every step."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    m = _mirror(dev)
    if m is None:
        return 0
    bits = sum((1 << k) for k in range(4) if m[k])
    if reset:
        for k in range(4):
            m[k] = 0
    return bits


def _raise_if_timed_out(dev: torch.device) -> None:
    """A kernel of an EARLIER call gave up waiting (MAS_STATUS_TIMEOUT): that call's affected utterances got an
    all-zero path.  Fail loudly now rather than let training continue on it."""
    m = _mirror(dev)
    if m is not None and m[3]:
        m[3] = 0
        raise _lib.MasError("an earlier maximum_path call on this device timed out waiting for one of its own kernels "
                            "(MAS_STATUS_TIMEOUT): its alignment is all-zero for the affected utterances and must not be used")


def _decode_status(bits: int) -> str:
    msgs = []
    if bits & _lib.MAS_STATUS_TX_GT_TY:
        msgs.append("an utterance has t_x > t_y (undefined in the reference, core.pyx:16,32)")
    if bits & _lib.MAS_STATUS_EMPTY:
        msgs.append("an utterance has t_x < 1 or t_y < 1")
    if bits & _lib.MAS_STATUS_TOO_LONG:
        msgs.append("an utterance is longer than the padded tensor")
    if bits & 8:
        msgs.append("internal: a kernel gave up waiting for another (results are invalid)")
    return "; ".join(msgs)


def last_status(device=None, reset: bool = True) -> int:
    """Sticky MAS_STATUS_* bits accumulated by calls on ``device``'s current stream (synchronises)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(dev).cuda_stream)
    buf = _scratch.get(key)
    if buf is None:
        return 0
    word = buf[:4].view(torch.int32)
    bits = int(word.item())
    if reset and bits:
        word.zero_()
    return bits


def _check_neg_cent(neg_cent: torch.Tensor) -> None:
    if neg_cent.dim() != 3:
        raise ValueError(f"neg_cent must be [b, t_t, t_s], got {tuple(neg_cent.shape)}")
    if not neg_cent.is_floating_point():
        raise ValueError(f"neg_cent must be a floating tensor, got {neg_cent.dtype}")
    if not neg_cent.is_contiguous():
        # the reference's typed memoryview float[:,:,::1] rejects this too (core.pyx:38)
        raise ValueError("ndarray is not C-contiguous")


def _run_cuda(neg_cent: torch.Tensor, *, mask: Optional[torch.Tensor], t_ys: Optional[torch.Tensor],
              t_xs: Optional[torch.Tensor], want_path: bool, want_index: bool, check: bool):
    L = _lib.lib()
    dev = neg_cent.device
    B, T_y, T_x = neg_cent.shape
    values = neg_cent.detach()
    if values.dtype != torch.float32:
        values = values.float()  # the reference always computes in float32 (__init__.py:14)
    out_dtype = neg_cent.dtype
    with torch.cuda.device(dev):
        _raise_if_timed_out(dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        nbytes = int(L.mas_maximum_path_scratch_bytes(B, T_y, T_x))
        scratch = _scratch_for(dev, stream, nbytes)
        path = torch.empty((B, T_y, T_x), dtype=out_dtype, device=dev) if want_path else None
        index = torch.empty((B, T_y), dtype=torch.int32, device=dev) if want_index else None
        if mask is not None:
            m = mask.detach()
            if m.device != dev:
                raise ValueError("mask must be on neg_cent's device")
            if m.dim() != 3 or tuple(m.shape) != (B, T_y, T_x):
                raise ValueError(f"mask must be [b, t_t, t_s] = {(B, T_y, T_x)}, got {tuple(m.shape)}")
            if m.dtype not in _DT:
                raise ValueError(f"unsupported mask dtype {m.dtype}")
            margs = (m.data_ptr(), _DT[m.dtype], m.stride(0), m.stride(1), m.stride(2))
            largs = (None, None)
        else:
            t_ys = t_ys.to(device=dev, dtype=torch.int32).contiguous()
            t_xs = t_xs.to(device=dev, dtype=torch.int32).contiguous()
            if t_ys.numel() != B or t_xs.numel() != B:
                raise ValueError("lengths must have one entry per utterance")
            margs = (None, 0, 0, 0, 0)
            largs = (t_ys.data_ptr(), t_xs.data_ptr())
        rc = L.mas_maximum_path(values.data_ptr(), largs[0], largs[1], *margs,
                                path.data_ptr() if want_path else None, _DT[out_dtype],
                                index.data_ptr() if want_index else None,
                                scratch.data_ptr(), scratch.numel(), B, T_y, T_x, stream)
        _lib.check(rc, "mas_maximum_path")
        if check or _CHECK:
            bits = last_status(dev)
            if bits:
                raise StatusError(_decode_status(bits))
    return path, index


def _lengths_from_mask_cpu(mask: torch.Tensor):
    # column 0 / row 0 sums, as __init__.py:17-18 (identical values, without reducing the full mask).  Through numpy where
    # numpy has the dtype: torch's multi-threaded reduction over the strided column-0 view costs 30-60 ms for a
    # [64,1024,192] mask on an 8-thread host (0.4 ms single-threaded, 0.4 ms in numpy) -- more than the whole alignment.
    m = mask.detach()
    try:
        a = m.numpy()
    except (TypeError, RuntimeError):   # bfloat16 and friends
        a = None
    if a is not None:
        t_ys = torch.from_numpy(a[:, :, 0].sum(1, dtype=np.float64).astype(np.int32))   # (exact counts, as the kernels')
        t_xs = torch.from_numpy(a[:, 0, :].sum(1, dtype=np.float64).astype(np.int32))
    else:
        t_ys = m[:, :, 0].float().sum(1).to(torch.int32)
        t_xs = m[:, 0, :].float().sum(1).to(torch.int32)
    return t_ys, t_xs


def _run_host(neg_cent: torch.Tensor, t_ys: torch.Tensor, t_xs: torch.Tensor, check: bool) -> torch.Tensor:
    """CPU tensors: the host-buffer C entry (H2D, kernels, D2H inside)."""
    L = _lib.lib()
    B, T_y, T_x = neg_cent.shape
    values = neg_cent.detach()
    if values.dtype != torch.float32:
        values = values.float()
    t_ys = t_ys.to(torch.int32).contiguous()
    t_xs = t_xs.to(torch.int32).contiguous()
    # the path leaves the device already in neg_cent's dtype (no int32 -> float pass on the host, __init__.py:20), and
    # the entry zeroes the padded rows it does not copy (no np.zeros pass, __init__.py:15)
    # Page-locked result (torch's caching host allocator: the cudaHostAlloc is paid once, later calls reuse the block):
    # the device writes it at link speed with no staging copy and no first-touch page faults; to the caller it is an
    # ordinary CPU tensor.
    try:
        paths = torch.empty((B, T_y, T_x), dtype=neg_cent.dtype, pin_memory=True)
    except RuntimeError:   # no usable CUDA driver: the C entry below reports that (there is no CPU implementation)
        paths = torch.empty((B, T_y, T_x), dtype=neg_cent.dtype)
    rc = L.mas_maximum_path_host(paths.data_ptr(), _DT[neg_cent.dtype], 1, values.data_ptr(), t_ys.data_ptr(),
                                 t_xs.data_ptr(), B, T_y, T_x)
    if rc > 0 and (rc & 0xFF) == 0:
        if check or _CHECK:
            raise StatusError(_decode_status(rc >> 8))
    else:
        _lib.check(rc, "mas_maximum_path_host")
    return paths


def maximum_path(neg_cent: torch.Tensor, mask: torch.Tensor, check: bool = False) -> torch.Tensor:
    """Monotonic alignment search.  Same call as the reference (monotonic_align/__init__.py:7).

    neg_cent: [b, t_t, t_s] float tensor (C-contiguous); mask: [b, t_t, t_s] prefix outer-product
    mask (float or bool).  Returns the 0/1 path, dtype/device of ``neg_cent``.
    ``check=True`` (or VITS_MAS_CHECK=1) synchronises and raises StatusError for lengths the
    reference leaves undefined (t_x > t_y, empty); otherwise such utterances get an all-zero path.
    """
    _check_neg_cent(neg_cent)
    if neg_cent.is_cuda:
        path, _ = _run_cuda(neg_cent, mask=mask, t_ys=None, t_xs=None, want_path=True, want_index=False, check=check)
        return path
    t_ys, t_xs = _lengths_from_mask_cpu(mask)
    return _run_host(neg_cent, t_ys, t_xs, check)


def maximum_path_from_lengths(neg_cent: torch.Tensor, y_lengths: torch.Tensor, x_lengths: torch.Tensor,
                              check: bool = False) -> torch.Tensor:
    """Same result as ``maximum_path`` with the lengths given directly (core.pyx:38 ``t_ys``,
    ``t_xs``) instead of through the dense mask: skips materialising ``attn_mask``."""
    _check_neg_cent(neg_cent)
    if neg_cent.is_cuda:
        path, _ = _run_cuda(neg_cent, mask=None, t_ys=y_lengths, t_xs=x_lengths, want_path=True, want_index=False,
                            check=check)
        return path
    return _run_host(neg_cent, y_lengths.cpu(), x_lengths.cpu(), check)


def maximum_path_index(neg_cent: torch.Tensor, mask: Optional[torch.Tensor] = None, *,
                       y_lengths: Optional[torch.Tensor] = None, x_lengths: Optional[torch.Tensor] = None,
                       check: bool = False) -> torch.Tensor:
    """Compact form of the path: int32 ``[B, T_y]`` text position per frame (-1 on padded
    frames).  ``path[b, y, index[b, y]] == 1``.  CUDA tensors only."""
    _check_neg_cent(neg_cent)
    if not neg_cent.is_cuda:
        raise ValueError("maximum_path_index needs CUDA tensors")
    if mask is None and (y_lengths is None or x_lengths is None):
        raise ValueError("give either mask or both y_lengths and x_lengths")
    _, index = _run_cuda(neg_cent, mask=mask, t_ys=y_lengths, t_xs=x_lengths, want_path=False, want_index=True,
                         check=check)
    return index
