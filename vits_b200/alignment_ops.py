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


class _KlIndex(torch.autograd.Function):
    @staticmethod
    def forward(ctx, index, z_p, logs_q, m_p, logs_p, z_mask):
        idx = _check_index(index)
        B, T_y = idx.shape
        z, lq, m, lp = (t.detach().float().contiguous() for t in (z_p, logs_q, m_p, logs_p))
        mk = z_mask.detach().float().reshape(B, T_y).contiguous()
        C, T_x = m.shape[1], m.shape[2]
        out = torch.empty(2, dtype=torch.float64, device=idx.device)
        with torch.cuda.device(idx.device):
            rc = _lib.lib().mas_kl_from_index(idx.data_ptr(), z.data_ptr(), lq.data_ptr(), m.data_ptr(), lp.data_ptr(),
                                              mk.data_ptr(), out.data_ptr(), B, C, T_y, T_x, _stream(idx.device))
        _lib.check(rc, "mas_kl_from_index")
        ctx.save_for_backward(idx, z, lq, m, lp, mk, out)
        return (out[0] / out[1]).float()

    @staticmethod
    def backward(ctx, g):
        idx, z, lq, m, lp, mk, out = ctx.saved_tensors
        scale = (g.double() / out[1]).float()
        # recompute on the expanded statistics (our gather), then scatter the text-side gradients back
        m_e, lp_e = _Expand.apply(idx, m, lp)
        w = mk[:, None, :] * scale
        iv = torch.exp(-2.0 * lp_e)
        d = z - m_e
        g_z = d * iv * w
        g_lq = -w.expand_as(z)
        g_me = -g_z
        g_lpe = (1.0 - d * d * iv) * w
        valid = idx >= 0
        at = idx.clamp_min(0).to(torch.int64)[:, None, :].expand_as(z)

        def scatter(t):
            t = t * valid[:, None, :]
            return torch.zeros_like(m).scatter_add_(2, at, t)

        return None, g_z, g_lq.contiguous(), scatter(g_me), scatter(g_lpe), None


def kl_loss_from_index(index: torch.Tensor, z_p: torch.Tensor, logs_q: torch.Tensor, m_p: torch.Tensor,
                       logs_p: torch.Tensor, z_mask: torch.Tensor) -> torch.Tensor:
    """The reference's ``kl_loss(z_p, logs_q, m_p, logs_p, z_mask)`` (losses.py:43-60) with the TEXT-side prior
    statistics ``m_p, logs_p [B, C, T_x]`` and the alignment ``index`` instead of the expanded ``[B, C, T_y]``
    tensors: the expansion (SynthesizerTrn.py:247-248) happens inside the reduction.  ``z_mask`` is ``[B, 1, T_y]``.
    Differentiable with respect to z_p, logs_q, m_p, logs_p."""
    return _KlIndex.apply(index, z_p, logs_q, m_p, logs_p, z_mask)


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
