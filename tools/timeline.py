"""Kernel timeline of one maximum_path call from %globaltimer stamps (mas_set_timeline).
usage: python tools/timeline.py [c2|c3|c4]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vits_b200
from vits_b200 import _lib
from bench import WORKLOADS, make_lengths

wl = sys.argv[1] if len(sys.argv) > 1 else 'c2'
B, T_y, T_x = WORKLOADS[wl]
L = _lib.lib()
t_ys, t_xs = make_lengths(np.random.default_rng(0), B, T_y, T_x, '--ragged' in sys.argv)
ty, tx = torch.as_tensor(t_ys).cuda(), torch.as_tensor(t_xs).cuda()
g = torch.Generator(device='cuda').manual_seed(1)
bufs = [torch.randn(B, T_y, T_x, generator=g, device='cuda') * 20 - 400 for _ in range(4)]
names = ['fwd first start', 'fwd last DP warp done', 'fwd last end', 'bt first start', 'bt last end',
         'wo first start', 'wo last zero-fill done', 'wo last end']
tl = torch.zeros(16, dtype=torch.int64, device='cuda')
def reset():
    tl.zero_(); tl[0] = tl[3] = tl[5] = -1   # uint64 max for the "min" slots
def show(tag):
    torch.cuda.synchronize()
    v = tl.cpu().numpy().astype(np.uint64)
    t0 = int(v[0])
    print(tag)
    for n, x in zip(names, v):
        x = int(x)
        if x in (0, 2**64 - 1): continue
        print(f"   {n:26s} {(x - t0) / 1e3:8.2f} us")
for mode in ('eager', 'graph'):
    for pdl in (1, 0):
        L.mas_set_tuning(0, 0, 0, pdl)
        for i in range(2): vits_b200.maximum_path_from_lengths(bufs[i], ty, tx)
        torch.cuda.synchronize()
        if mode == 'eager':
            reset(); L.mas_set_timeline(tl.data_ptr())
            vits_b200.maximum_path_from_lengths(bufs[2], ty, tx)
            show(f"{wl} eager pdl={pdl}")
            L.mas_set_timeline(None)
        else:
            L.mas_set_timeline(tl.data_ptr())
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                out = vits_b200.maximum_path_from_lengths(bufs[3], ty, tx)
            gr.replay(); torch.cuda.synchronize()
            reset(); gr.replay()
            show(f"{wl} graph pdl={pdl}")
            L.mas_set_timeline(None)
L.mas_set_tuning(0, 0, 0, -1)
