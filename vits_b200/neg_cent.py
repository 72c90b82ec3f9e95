"""``neg_cent(z_p, m_p, logs_p)``: the contraction SynthesizerTrn.forward computes inline
(reference SynthesizerTrn.py:223-232), as one fused sm_100a kernel behind ``mas_neg_cent``
(include/vits_mas.h).  fp32 in, fp32 out, independent of any surrounding autocast context."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _lib

_scratch: Dict[Tuple[int, int], torch.Tensor] = {}


def _scratch_for(device: torch.device, stream: int, nbytes: int) -> torch.Tensor:
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


def neg_cent(z_p: torch.Tensor, m_p: torch.Tensor, logs_p: torch.Tensor, *,
             autocast_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """z_p [B,C,T_y]; m_p, logs_p [B,C,T_x]  ->  neg_cent [B,T_y,T_x] float32.

    Inputs of other float dtypes (e.g. fp16 activations under autocast) are promoted to float32
    first: the parity target is the reference's fp32 formulation.

    ``autocast_dtype=torch.float16`` (or ``bfloat16``) selects the autocast-parity mode instead: the numerics
    the reference has as trained under ``autocast(fp16_run)`` (train_and_evaluate.py:55) -- both einsums on
    operands rounded to that type with their outputs rounded to it, everything else fp32 -- except that when
    ``logs_p`` itself arrives in that type (TextEncoder.proj's output under autocast) the elementwise part of
    :225 is rounded to it too, as plain tensor arithmetic in that dtype is.  It runs on CUDA cores and is there
    to reproduce as-trained alignments, not for speed."""
    if not (z_p.is_cuda and m_p.is_cuda and logs_p.is_cuda):
        raise ValueError("neg_cent needs CUDA tensors (there is no CPU implementation)")
    if z_p.dim() != 3 or m_p.dim() != 3 or m_p.shape != logs_p.shape or z_p.shape[:2] != m_p.shape[:2]:
        raise ValueError(f"bad shapes z_p {tuple(z_p.shape)} m_p {tuple(m_p.shape)} logs_p {tuple(logs_p.shape)}")
    L = _lib.lib()
    z, m, ls = (t.detach().float().contiguous() for t in (z_p, m_p, logs_p))
    B, C, T_y = z.shape
    T_x = m.shape[2]
    dev = z.device
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        out = torch.empty((B, T_y, T_x), dtype=torch.float32, device=dev)
        if autocast_dtype is not None:
            code = {torch.float16: 1, torch.bfloat16: 2}.get(autocast_dtype)
            if code is None:
                raise ValueError("autocast_dtype must be torch.float16 or torch.bfloat16")
            stats_lowp = int(logs_p.dtype == autocast_dtype)   # :225 ran in logs_p's own dtype
            rc = L.mas_neg_cent_autocast(z.data_ptr(), m.data_ptr(), ls.data_ptr(), out.data_ptr(), code, stats_lowp,
                                         B, C, T_y, T_x, stream)
            _lib.check(rc, "mas_neg_cent_autocast")
            return out
        nbytes = int(L.mas_neg_cent_scratch_bytes(B, C, T_y, T_x))
        scratch = _scratch_for(dev, stream, nbytes)
        rc = L.mas_neg_cent(z.data_ptr(), m.data_ptr(), ls.data_ptr(), out.data_ptr(), scratch.data_ptr(),
                            scratch.numel(), B, C, T_y, T_x, stream)
        _lib.check(rc, "mas_neg_cent")
    return out


def maximum_path_from_stats(z_p: torch.Tensor, m_p: torch.Tensor, logs_p: torch.Tensor, x_lengths: torch.Tensor,
                            y_lengths: torch.Tensor, *, index: bool = False) -> torch.Tensor:
    """``z_p, m_p, logs_p -> path`` in one call: the contraction (SynthesizerTrn.py:223-232) followed by the
    alignment search (:235), with the lengths given directly -- no ``[B, T_y, T_x]`` mask is built or read
    (SynthesizerTrn.py:234).  ``index=True`` returns the compact int32 ``[B, T_y]`` form instead of the dense
    path.  (The two stages still exchange ``neg_cent`` through HBM; SURVEY.md 8f rank 1 would fuse them.)"""
    from .monotonic_align import maximum_path_from_lengths, maximum_path_index
    nc = neg_cent(z_p, m_p, logs_p)
    if index:
        return maximum_path_index(nc, y_lengths=y_lengths, x_lengths=x_lengths)
    return maximum_path_from_lengths(nc, y_lengths, x_lengths)
