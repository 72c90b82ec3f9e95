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


_fused_scratch: Dict[Tuple[int, int], torch.Tensor] = {}
_STREAMED_MAX_TILES = 16   # [128 frames x <= 256 tokens] contraction tiles up to which the streamed form is the faster one


def maximum_path_from_stats(z_p: torch.Tensor, m_p: torch.Tensor, logs_p: torch.Tensor, x_lengths: torch.Tensor,
                            y_lengths: torch.Tensor, *, index: bool = False, streamed: Optional[bool] = None):
    """``z_p, m_p, logs_p -> path`` in one call: the contraction (SynthesizerTrn.py:223-232) and the alignment search
    (:235), with the lengths given directly -- no ``[B, T_y, T_x]`` mask is built or read (SynthesizerTrn.py:234).

    ``streamed=True`` runs the two stages STREAMED (C entry ``mas_stats_to_path``): the tensor-core contraction on one
    part of the SMs feeds the search on the others tile by tile through an L2-resident ring, so ``neg_cent`` never makes
    the HBM round trip and the stages overlap; paths are bit-identical to ``maximum_path(neg_cent(...), mask)``.
    ``streamed=False`` runs the two kernels back to back (through HBM).  The default picks by measurement (DESIGN.md
    section 4.6): streamed wins while the contraction is a single wave of tiles (small batches: 29 vs 39 us at B = 3),
    the two-call form wins at training-size batches, where the contraction needs every SM (c2: 86 vs 102 us).  Returns the dense path ``[B, T_y, T_x]`` float32; ``index=True`` returns the compact
    int32 ``[B, T_y]`` form instead (text position per frame, -1 on padded frames)."""
    from .monotonic_align import _raise_if_timed_out, maximum_path_from_lengths, maximum_path_index
    if not (z_p.is_cuda and m_p.is_cuda and logs_p.is_cuda):
        raise ValueError("maximum_path_from_stats needs CUDA tensors (there is no CPU implementation)")
    if z_p.dim() != 3 or m_p.dim() != 3 or m_p.shape != logs_p.shape or z_p.shape[:2] != m_p.shape[:2]:
        raise ValueError(f"bad shapes z_p {tuple(z_p.shape)} m_p {tuple(m_p.shape)} logs_p {tuple(logs_p.shape)}")
    L = _lib.lib()
    B, C, T_y = z_p.shape
    T_x = m_p.shape[2]
    dev = z_p.device
    if streamed is None:
        streamed = B * (-(-T_y // 128)) * (-(-T_x // 256)) <= _STREAMED_MAX_TILES and T_x <= 512
    if streamed:
        z, m, ls = (t.detach().float().contiguous() for t in (z_p, m_p, logs_p))
        t_ys = y_lengths.to(device=dev, dtype=torch.int32).contiguous()
        t_xs = x_lengths.to(device=dev, dtype=torch.int32).contiguous()
        if t_ys.numel() != B or t_xs.numel() != B:
            raise ValueError("lengths must have one entry per utterance")
        with torch.cuda.device(dev):
            _raise_if_timed_out(dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            nbytes = int(L.mas_stats_to_path_scratch_bytes(B, C, T_y, T_x))
            key = (dev.index if dev.index is not None else torch.cuda.current_device(), stream)
            buf = _fused_scratch.get(key)
            if buf is None or buf.numel() < nbytes:
                buf = torch.zeros(max(nbytes, 1 << 16), dtype=torch.uint8, device=dev)   # (sticky status word inside)
                _fused_scratch[key] = buf
            path = torch.empty((B, T_y, T_x), dtype=torch.float32, device=dev)
            idx = torch.empty((B, T_y), dtype=torch.int32, device=dev) if index else None
            rc = L.mas_stats_to_path(z.data_ptr(), m.data_ptr(), ls.data_ptr(), t_ys.data_ptr(), t_xs.data_ptr(),
                                     path.data_ptr(), _lib.MAS_F32, idx.data_ptr() if index else None, buf.data_ptr(),
                                     buf.numel(), B, C, T_y, T_x, stream)
        if rc == 0:
            return idx if index else path
        _lib.check(rc, "mas_stats_to_path")
    nc = neg_cent(z_p, m_p, logs_p)
    if index:
        return maximum_path_index(nc, y_lengths=y_lengths, x_lengths=x_lengths)
    return maximum_path_from_lengths(nc, y_lengths, x_lengths)
