"""Multi-rank host logic on CPU (gloo, world_size 2): shard bounds, index all-gather with uneven shards,
and the verification the benchmark runs after its timed region.  The per-shard compute is stood in for by
the oracle (this is a CPU test; the CUDA path has no CPU fallback)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, T_y, T_x, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import path_to_index, random_lengths
        from oracle import mas_oracle
        from vits_b200 import shard
        rng = np.random.default_rng(99)                       # same batch on every rank
        nc = (rng.standard_normal((B, T_y, T_x)) * 2).astype(np.float32)
        t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
        lo, hi = shard.shard_bounds(B, rank, world)
        mine = path_to_index(mas_oracle.maximum_path_numpy(nc[lo:hi], t_ys[lo:hi], t_xs[lo:hi]))
        gathered = shard.gather_index(torch.from_numpy(mine), B)
        ok = True
        if rank == 0:
            want = path_to_index(mas_oracle.maximum_path_numpy(nc, t_ys, t_xs))
            ok = bool(np.array_equal(gathered.numpy(), want)) and shard.check_index(gathered, t_ys, t_xs)
            bad = gathered.clone()
            bad[1, 3] += 2                                     # a corrupted shard must be caught
            ok = ok and not shard.check_index(bad, t_ys, t_xs)
        q.put((rank, lo, hi, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [6, 7])
def test_sharded_alignment_gathers_to_the_single_rank_result(B):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, 40, 12, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[3] for r in res)
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == B     # contiguous cover


def test_shard_bounds_cover_and_balance():
    from vits_b200 import shard
    for n in (1, 5, 64, 65, 513):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_bounds(4, 2, 2)
