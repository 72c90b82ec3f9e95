"""Time mas_maximum_path_c_host on pinned buffers (c2, both length variants) -- run under different MAS_HOST_GROUPS."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vits_b200
from vits_b200 import _lib
from bench import WORKLOADS, make_lengths
L = _lib.lib()
B, T_y, T_x = WORKLOADS["c2"]
for ragged in (True, False):
    rng = np.random.default_rng(1234)
    t_ys, t_xs = make_lengths(rng, B, T_y, T_x, ragged)
    vals = torch.randn(B, T_y, T_x).pin_memory()
    paths = torch.zeros(B, T_y, T_x, dtype=torch.int32).pin_memory()
    ty, tx = torch.as_tensor(t_ys, dtype=torch.int32), torch.as_tensor(t_xs, dtype=torch.int32)
    def step():
        rc = L.mas_maximum_path_c_host(paths.data_ptr(), vals.data_ptr(), ty.data_ptr(), tx.data_ptr(), B, T_y, T_x)
        assert rc == 0, rc
    for _ in range(3): step()
    t0 = time.perf_counter()
    for _ in range(20): step()
    dt = (time.perf_counter() - t0) / 20
    print(f"groups={os.environ.get('MAS_HOST_GROUPS','default')} ragged={ragged}: {dt*1e3:.3f} ms/call -> {B/dt:,.0f} alignments/s")
if os.environ.get("MAS_LINK"):
    d_a = torch.empty(B, T_y, T_x, device="cuda"); d_b = torch.randn(B, T_y, T_x, device="cuda")
    h_b = torch.empty(B, T_y, T_x).pin_memory()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def both(n):
        for _ in range(n):
            with torch.cuda.stream(s1): d_a.copy_(vals, non_blocking=True)
            with torch.cuda.stream(s2): h_b.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
    both(2)
    t0 = time.perf_counter(); both(10); dt = (time.perf_counter() - t0) / 10
    nbytes = vals.numel() * 4
    print(f"duplex: {2*nbytes/dt/1e9:.1f} GB/s combined ({dt*1e3:.3f} ms for 2 x {nbytes/1e6:.1f} MB) -> bound {B/dt:,.0f} alignments/s full-length")
