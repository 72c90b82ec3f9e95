"""Minimal driver for ncu: streamed stats->path calls and two-call calls on one workload (default c2 full-length)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vits_b200
from bench import WORKLOADS, make_lengths
wl = sys.argv[1] if len(sys.argv) > 1 else 'c2'
B, T_y, T_x = WORKLOADS[wl]
t_ys, t_xs = make_lengths(np.random.default_rng(1234), B, T_y, T_x, '--ragged' in sys.argv)
ty, tx = torch.as_tensor(t_ys).cuda(), torch.as_tensor(t_xs).cuda()
g = torch.Generator(device='cuda').manual_seed(1)
sets = [(torch.randn(B, 192, T_y, generator=g, device='cuda'), torch.randn(B, 192, T_x, generator=g, device='cuda'),
         torch.randn(B, 192, T_x, generator=g, device='cuda') * 0.3) for _ in range(2)]
for i in range(2):
    a = vits_b200.maximum_path_from_stats(*sets[i], tx, ty, streamed=True)
for i in range(2):
    b = vits_b200.maximum_path_from_stats(*sets[i], tx, ty, streamed=False)
torch.cuda.synchronize()
print("ok", float(a.sum()), float(b.sum()), bool(torch.equal(a, b)))
