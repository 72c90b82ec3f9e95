"""Timing of the path's consumers at a workload (GPU box): ours on the compact index vs the reference's torch expressions."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import vits_b200
from oracle import mas_oracle
from bench import WORKLOADS
wl = sys.argv[1] if len(sys.argv) > 1 else 'c2'
B, T_y, T_x = WORKLOADS[wl]; C = 192
g = torch.Generator(device='cuda').manual_seed(1)
nc = torch.randn(B, T_y, T_x, generator=g, device='cuda') * 20 - 400
ty = torch.full((B,), T_y, device='cuda'); tx = torch.full((B,), T_x, device='cuda')
index = vits_b200.maximum_path_index(nc, y_lengths=ty, x_lengths=tx)
attn = vits_b200.maximum_path_from_lengths(nc, ty, tx).unsqueeze(1)
m_p = torch.randn(B, C, T_x, generator=g, device='cuda'); logs_p = torch.randn(B, C, T_x, generator=g, device='cuda') * 0.3
z_p = torch.randn(B, C, T_y, generator=g, device='cuda'); logs_q = torch.randn(B, C, T_y, generator=g, device='cuda') * 0.3
z_mask = torch.ones(B, 1, T_y, device='cuda')
dur = vits_b200.path_durations(index, T_x); mask4 = torch.ones(B, 1, T_y, T_x, device='cuda')
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
with torch.no_grad():
    rows = [
        ("durations  w = attn.sum(2)", lambda: vits_b200.path_durations(index, T_x), lambda: mas_oracle.durations_torch(attn)),
        ("prior expansion (m_p and logs_p)", lambda: vits_b200.expand_prior(index, m_p, logs_p),
         lambda: (mas_oracle.expand_prior_torch(attn, m_p), mas_oracle.expand_prior_torch(attn, logs_p))),
        ("kl_loss incl. the expansion", lambda: vits_b200.kl_loss_from_index(index, z_p, logs_q, m_p, logs_p, z_mask),
         lambda: mas_oracle.kl_loss_torch(z_p, logs_q, mas_oracle.expand_prior_torch(attn, m_p), mas_oracle.expand_prior_torch(attn, logs_p), z_mask)),
        ("generate_path(duration, mask)", lambda: vits_b200.generate_path(dur, mask4), lambda: mas_oracle.generate_path_torch(dur, mask4)),
    ]
    print(f"workload {wl}: B={B} T_y={T_y} T_x={T_x} C={C} (eager launches, CUDA events)")
    for name, ours, ref in rows:
        print(f"{name:36s} ours {timeit(ours):8.1f} us | reference torch expression {timeit(ref):8.1f} us")
