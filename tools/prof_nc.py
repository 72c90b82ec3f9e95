"""Minimal driver for ncu: a few neg_cent calls at one workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vits_b200
from bench import WORKLOADS
wl = sys.argv[1] if len(sys.argv) > 1 else 'c2'
B, T_y, T_x = WORKLOADS[wl]; C = 192
g = torch.Generator(device='cuda').manual_seed(1)
z = torch.randn(B, C, T_y, generator=g, device='cuda'); m = torch.randn(B, C, T_x, generator=g, device='cuda')
ls = torch.randn(B, C, T_x, generator=g, device='cuda') * 0.3
for i in range(3):
    out = vits_b200.neg_cent(z, m, ls)
torch.cuda.synchronize()
print("ok", float(out[0, 0, 0]))
