"""Callers either side of the alignment path, on the compact per-frame index.

    path_durations(index, T_x)           == attn.sum(2)                               SynthesizerTrn.py:237
    expand_prior(index, m_p, logs_p)     == einsum('bctn,bdn->bdt', attn, m_p|logs_p) SynthesizerTrn.py:247-248
    generate_path(duration, mask)        == commons.generate_path                     commons.py:101-117

``index`` is the int32 ``[B, T_y]`` tensor of ``maximum_path_index`` (text position per frame, -1 on padded
frames): with it the sum over frames is a histogram and the one-hot GEMMs are gathers, and the dense
``[B, T_y, T_x]`` path is never read back.  CUDA tensors only (there is no CPU implementation).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _check_index(index: torch.Tensor) -> torch.Tensor:
    if not index.is_cuda or index.dim() != 2:
        raise ValueError("index must be a CUDA tensor [B, T_y]")
    return index.to(torch.int32).contiguous()


def path_durations(index: torch.Tensor, T_x: int) -> torch.Tensor:
    """Frames per text position: float32 ``[B, 1, T_x]`` like ``attn.sum(2)`` of the reference's ``[B,1,T_y,T_x]`` path."""
    idx = _check_index(index)
    B, T_y = idx.shape
    w = torch.empty((B, 1, T_x), dtype=torch.float32, device=idx.device)
    with torch.cuda.device(idx.device):
        rc = _lib.lib().mas_path_durations(idx.data_ptr(), w.data_ptr(), B, T_y, T_x, _stream(idx.device))
    _lib.check(rc, "mas_path_durations")
    return w


class _Expand(torch.autograd.Function):
    @staticmethod
    def forward(ctx, index, m_p, logs_p):
        idx = _check_index(index)
        B, T_y = idx.shape
        srcs = [t.detach().float().contiguous() for t in (m_p, logs_p) if t is not None]
        C, T_x = srcs[0].shape[1], srcs[0].shape[2]
        outs = [torch.empty((B, C, T_y), dtype=torch.float32, device=idx.device) for _ in srcs]
        with torch.cuda.device(idx.device):
            rc = _lib.lib().mas_expand_prior(idx.data_ptr(), srcs[0].data_ptr(), srcs[1].data_ptr() if len(srcs) > 1 else None,
                                             outs[0].data_ptr(), outs[1].data_ptr() if len(srcs) > 1 else None,
                                             B, C, T_y, T_x, _stream(idx.device))
        _lib.check(rc, "mas_expand_prior")
        ctx.save_for_backward(idx)
        ctx.T_x = T_x
        ctx.two = logs_p is not None
        return (outs[0], outs[1]) if ctx.two else (outs[0], None)

    @staticmethod
    def backward(ctx, g_m, g_l):
        # grad_src[b,c,x] = sum of grad_out[b,c,y] over the frames aligned to x (the transpose of the gather)
        (idx,) = ctx.saved_tensors
        valid = idx >= 0
        gather_at = idx.clamp_min(0).to(torch.int64)

        def scatter(g):
            if g is None:
                return None
            g = g.float() * valid[:, None, :]
            out = torch.zeros(g.shape[0], g.shape[1], ctx.T_x, dtype=torch.float32, device=g.device)
            return out.scatter_add_(2, gather_at[:, None, :].expand_as(g), g)

        return None, scatter(g_m), scatter(g_l) if ctx.two else None


def expand_prior(index: torch.Tensor, m_p: torch.Tensor, logs_p: Optional[torch.Tensor] = None
                 ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Expand the text-side prior statistics ``[B, C, T_x]`` to frames ``[B, C, T_y]`` along the alignment.
    Differentiable with respect to ``m_p`` and ``logs_p``."""
    if m_p.dim() != 3 or (logs_p is not None and logs_p.shape != m_p.shape):
        raise ValueError("m_p / logs_p must be [B, C, T_x]")
    return _Expand.apply(index, m_p, logs_p)


def generate_path(duration: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """Drop-in for the reference's ``commons.generate_path(duration, mask)``: duration ``[b, 1, t_x]``,
    mask ``[b, 1, t_y, t_x]``; returns the ``[b, 1, t_y, t_x]`` path in ``mask``'s dtype."""
    if not (duration.is_cuda and mask.is_cuda):
        raise ValueError("generate_path needs CUDA tensors (there is no CPU implementation)")
    b, _, t_y, t_x = mask.shape
    d = duration.detach().reshape(b, t_x).float().contiguous()
    m = mask.detach().reshape(b, t_y, t_x)
    if m.dtype != torch.float32:
        m = m.float()
    out = torch.empty((b, t_y, t_x), dtype=torch.float32, device=mask.device)
    with torch.cuda.device(mask.device):
        rc = _lib.lib().mas_generate_path(d.data_ptr(), m.data_ptr(), m.stride(0), m.stride(1), m.stride(2), out.data_ptr(),
                                          b, t_y, t_x, _stream(mask.device))
    _lib.check(rc, "mas_generate_path")
    return out.to(mask.dtype).unsqueeze(1)
