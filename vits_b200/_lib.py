"""ctypes binding of libvits_mas.so (the C ABI in include/vits_mas.h).

There is no CPU fallback: if the library cannot be built or loaded, importing the compute
entry points raises.  Nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes
import os

from . import _build

_lib = None

# element-type codes (include/vits_mas.h)
MAS_F32, MAS_F16, MAS_BF16, MAS_F64, MAS_U8, MAS_I8, MAS_I16, MAS_I32, MAS_I64 = range(9)
MAS_STATUS_TX_GT_TY, MAS_STATUS_EMPTY, MAS_STATUS_TOO_LONG, MAS_STATUS_TIMEOUT = 1, 2, 4, 8
ABI_VERSION = 2

EXPORTS = [
    "mas_abi_version", "mas_status_mirror", "mas_last_error_site", "mas_error_string", "mas_maximum_path_scratch_bytes", "mas_scratch_status_offset",
    "mas_maximum_path", "mas_maximum_path_c_host", "mas_maximum_path_host", "mas_host_release", "mas_neg_cent_scratch_bytes",
    "mas_neg_cent", "mas_stats_to_path_scratch_bytes", "mas_stats_to_path", "mas_neg_cent_autocast", "mas_path_durations", "mas_expand_prior", "mas_generate_path", "mas_kl_from_index", "mas_launch_count", "mas_set_tuning", "mas_set_neg_cent_impl", "mas_set_debug_kernels", "mas_set_tuning2", "mas_set_tuning3", "mas_set_timeline", "mas_set_trace",
]


class MasError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    # (re)build when the library is missing or stale against csrc/ (content hash; a no-op otherwise) -- an old
    # binary behind new ctypes signatures would fail in ways that are hard to read.  Raises when nvcc is missing.
    path = _build.build()
    if os.environ.get("VITS_MAS_LIB"):      # experiments: an alternative build of the same ABI (tools/build_variant.py)
        path = os.environ["VITS_MAS_LIB"]
    L = ctypes.CDLL(path)
    c_int, c_i64, c_sz, c_vp = ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_void_p
    L.mas_abi_version.restype = c_int
    if L.mas_abi_version() != ABI_VERSION:
        raise MasError(f"libvits_mas.so has ABI {L.mas_abi_version()}, this binding expects {ABI_VERSION}: rebuild "
                       "(python -c 'import __graft_entry__ as g; g.build()')")
    L.mas_last_error_site.restype = c_int
    L.mas_status_mirror.restype = ctypes.POINTER(ctypes.c_int32)
    L.mas_status_mirror.argtypes = []
    L.mas_error_string.restype = ctypes.c_char_p
    L.mas_error_string.argtypes = [c_int]
    L.mas_maximum_path_scratch_bytes.restype = c_sz
    L.mas_maximum_path_scratch_bytes.argtypes = [c_int, c_int, c_int]
    L.mas_scratch_status_offset.restype = c_sz
    L.mas_maximum_path.restype = c_int
    L.mas_maximum_path.argtypes = [c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_i64, c_i64, c_vp, c_int, c_vp, c_vp,
                                   c_sz, c_int, c_int, c_int, c_vp]
    L.mas_maximum_path_c_host.restype = c_int
    L.mas_maximum_path_c_host.argtypes = [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int]
    L.mas_maximum_path_host.restype = c_int
    L.mas_maximum_path_host.argtypes = [c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_int]
    L.mas_host_release.restype = None
    L.mas_neg_cent_scratch_bytes.restype = c_sz
    L.mas_neg_cent_scratch_bytes.argtypes = [c_int, c_int, c_int, c_int]
    L.mas_neg_cent.restype = c_int
    L.mas_neg_cent.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_int, c_int, c_int, c_int, c_vp]
    L.mas_stats_to_path_scratch_bytes.restype = c_sz
    L.mas_stats_to_path_scratch_bytes.argtypes = [c_int, c_int, c_int, c_int]
    L.mas_stats_to_path.restype = c_int
    L.mas_stats_to_path.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_sz, c_int, c_int, c_int, c_int, c_vp]
    L.mas_path_durations.restype = c_int
    L.mas_path_durations.argtypes = [c_vp, c_vp, c_int, c_int, c_int, c_vp]
    L.mas_expand_prior.restype = c_int
    L.mas_expand_prior.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp]
    L.mas_generate_path.restype = c_int
    L.mas_generate_path.argtypes = [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_int, c_int, c_int, c_vp]
    L.mas_kl_from_index.restype = c_int
    L.mas_kl_from_index.argtypes = [c_vp] * 7 + [c_int] * 4 + [c_vp]
    L.mas_launch_count.restype = ctypes.c_uint64
    L.mas_set_tuning.restype = None
    L.mas_set_tuning.argtypes = [c_int, c_int, c_int, c_int]
    L.mas_neg_cent_autocast.restype = c_int
    L.mas_neg_cent_autocast.argtypes = [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]
    L.mas_set_neg_cent_impl.restype = None
    L.mas_set_neg_cent_impl.argtypes = [c_int]
    L.mas_set_debug_kernels.restype = None
    L.mas_set_debug_kernels.argtypes = [c_int]
    L.mas_set_tuning2.restype = None
    L.mas_set_tuning2.argtypes = [c_int, c_int]
    L.mas_set_tuning3.restype = None
    L.mas_set_tuning3.argtypes = [c_int, c_int, c_int, c_int]
    L.mas_set_timeline.restype = None
    L.mas_set_timeline.argtypes = [c_vp]
    L.mas_set_trace.restype = None
    L.mas_set_trace.argtypes = [c_vp]
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().mas_error_string(rc).decode()
        site = lib().mas_last_error_site() if rc > 0 else 0
        raise MasError(f"{what} failed: {msg} (code {rc}" + (f", mas_path.cu:{site})" if site else ")"))


def launch_count() -> int:
    return int(lib().mas_launch_count())
