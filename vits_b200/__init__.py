"""vits_b200 -- B200 (sm_100a) implementation of the training-time alignment hot path of
Aloento/VITS: ``neg_cent`` (SynthesizerTrn.py:223-232) and ``monotonic_align.maximum_path``
(monotonic_align/__init__.py:7-20, core.pyx:7-42), behind the C ABI of include/vits_mas.h.

Public surface:
    vits_b200.monotonic_align.maximum_path(neg_cent, mask)     -- the reference's call, unchanged
    vits_b200.monotonic_align.maximum_path_from_lengths(...)   -- same, lengths instead of mask
    vits_b200.monotonic_align.maximum_path_index(...)          -- compact per-frame index
    vits_b200.neg_cent(z_p, m_p, logs_p)                       -- the contraction feeding it
    vits_b200.maximum_path_from_stats(z_p, m_p, logs_p, x_len, y_len)  -- both, without a mask tensor
    vits_b200.path_durations / expand_prior / generate_path    -- the path's consumers on the compact index, and
                                                                  commons.generate_path
    vits_b200.shard                                            -- batch sharding across GPUs (no data-path collective)
"""
from . import _lib  # noqa: F401
from . import monotonic_align  # noqa: F401
from .monotonic_align import (last_status, maximum_path, maximum_path_from_lengths, maximum_path_index,  # noqa: F401
                              status_nosync)
from .neg_cent import maximum_path_from_stats, neg_cent  # noqa: F401
from . import shard  # noqa: F401
from .alignment_ops import expand_prior, generate_path, kl_loss_from_index, path_durations  # noqa: F401

__version__ = "0.1.0"
