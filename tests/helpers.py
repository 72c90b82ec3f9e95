"""Shared test helpers: synthetic inputs and path property checks."""
import hashlib

import numpy as np


def index_to_path(index: np.ndarray, T_x: int) -> np.ndarray:
    B, T_y = index.shape
    path = np.zeros((B, T_y, T_x), dtype=np.int8)
    b, y = np.nonzero(index >= 0)
    path[b, y, index[b, y]] = 1
    return path


def path_to_index(path: np.ndarray) -> np.ndarray:
    p = np.asarray(path)
    idx = p.argmax(-1).astype(np.int32)
    idx[p.sum(-1) == 0] = -1
    return idx


def sha_path(path: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(path).astype(np.int8).tobytes()).hexdigest()


def random_lengths(rng, B, T_y, T_x, full_first=True):
    """t_x ~ U{ceil(T_x/2)..T_x}, t_y ~ U{max(t_x, ceil(T_y/2))..T_y}, element 0 full, sorted by t_y desc."""
    t_xs = rng.integers((T_x + 1) // 2, T_x + 1, size=B)
    t_ys = np.array([rng.integers(max(tx, (T_y + 1) // 2), T_y + 1) for tx in t_xs])
    if full_first:
        t_xs[0], t_ys[0] = T_x, T_y
    order = np.argsort(-t_ys, kind="stable")
    return t_ys[order].astype(np.int32), t_xs[order].astype(np.int32)


def check_path_properties(path: np.ndarray, t_ys, t_xs):
    """Size-independent invariants of a monotonic alignment."""
    path = np.asarray(path)
    B, T_y, T_x = path.shape
    assert set(np.unique(path)).issubset({0, 1})
    for b in range(B):
        ty, tx = int(t_ys[b]), int(t_xs[b])
        p = path[b]
        assert p[ty:].sum() == 0 and p[:, tx:].sum() == 0, "ones outside the valid box"
        rows = p[:ty, :tx]
        assert (rows.sum(1) == 1).all(), "every valid frame has exactly one 1"
        idx = rows.argmax(1)
        assert idx[0] == 0 and idx[-1] == tx - 1, "path must run from (0,0) to (t_y-1,t_x-1)"
        step = np.diff(idx)
        assert ((step == 0) | (step == 1)).all(), "index must be non-decreasing with steps in {0,1}"
    assert path.sum() == int(np.sum(t_ys))
