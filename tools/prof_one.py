"""Minimal driver for ncu: a few maximum_path calls on one workload (default c2 full-length)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vits_b200
from bench import WORKLOADS, make_lengths
wl = sys.argv[1] if len(sys.argv) > 1 else 'c2'
B, T_y, T_x = WORKLOADS[wl]
t_ys, t_xs = make_lengths(np.random.default_rng(1234), B, T_y, T_x, '--ragged' in sys.argv)  # bench.py's rank-0 lengths
ty, tx = torch.as_tensor(t_ys).cuda(), torch.as_tensor(t_xs).cuda()
g = torch.Generator(device='cuda').manual_seed(1)
bufs = [torch.randn(B, T_y, T_x, generator=g, device='cuda') * 20 - 400 for _ in range(3)]
for i in range(3):
    out = vits_b200.maximum_path_from_lengths(bufs[i], ty, tx)
torch.cuda.synchronize()
print("ok", float(out.sum()))
