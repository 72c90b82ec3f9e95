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


def near_tie_report(oracle, nc_ref: np.ndarray, nc_got: np.ndarray, idx_ref: np.ndarray, idx_got: np.ndarray, t_ys, t_xs):
    """SURVEY.md 8(d) gate 3: every frame where two alignments part must be a near-tie in the ORACLE's value table.

    ``nc_ref`` / ``nc_got`` are the two neg_cent tensors the paths were searched on, ``idx_*`` the per-frame text
    positions.  The backtrack (core.pyx:30-33) walks y = t_y-1 .. 0; the walks part when leaving frame y+1 at a common
    column i: the reference compares value[y, i] with value[y, i-1] (core.pyx:32).  A perturbation of at most
    ``e = max|nc_got - nc_ref|`` per cell moves the value of ANY path through y+1 rows by at most (y+1)*e, plus half an
    ulp of |value| per fp32 accumulation, so the decision can only flip when

        |value[y, i] - value[y, i-1]|  <=  2 (y+1) (e + ulp(max|value|)/2).

    Returns a list of (b, y, i, gap, bound); raises AssertionError on a divergence that is not such a near-tie or that
    is forced (column 0 / the diagonal, where core.pyx:32 does not compare at all)."""
    values = np.array(nc_ref, dtype=np.float32, order="C", copy=True)
    paths = np.zeros(values.shape, np.int32)
    oracle.maximum_path_c(paths, values, np.ascontiguousarray(t_ys, np.int32), np.ascontiguousarray(t_xs, np.int32))
    out = []
    B = values.shape[0]
    for b in range(B):
        ty, tx = int(t_ys[b]), int(t_xs[b])
        if ty < 1:
            continue
        e = float(np.abs(nc_got[b, :ty, :tx].astype(np.float64) - nc_ref[b, :ty, :tx]).max())
        v = values[b]
        ulp = float(np.spacing(np.float32(np.abs(v[:ty, :tx]).max())))
        r, g = idx_ref[b, :ty], idx_got[b, :ty]
        assert r[ty - 1] == g[ty - 1] == tx - 1
        for y in range(ty - 2, -1, -1):
            if r[y + 1] == g[y + 1] and r[y] != g[y]:
                i = int(r[y + 1])
                assert i != 0 and i != y + 1, f"utterance {b}: forced move differs at frame {y + 1}, column {i}"
                gap = abs(float(v[y, i]) - float(v[y, i - 1]))
                bound = 2.0 * (y + 1) * (e + 0.5 * ulp)
                assert gap <= bound, (f"utterance {b}: paths part at frame {y}, column {i} with |v_cur - v_prev| = {gap:.4g} "
                                      f"> near-tie bound {bound:.4g} (per-cell input error {e:.3g})")
                out.append((b, y, i, gap, bound))
    return out
