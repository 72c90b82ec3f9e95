"""Per-stage event trace of the forward kernel's CTA 0 (mas_set_trace).  usage: trace_fwd.py [c2] [fused 0/1]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vits_b200
from vits_b200 import _lib
from bench import WORKLOADS, make_lengths
wl = sys.argv[1] if len(sys.argv) > 1 else 'c2'
fused = int(sys.argv[2]) if len(sys.argv) > 2 else -1
B, T_y, T_x = WORKLOADS[wl]
L = _lib.lib()
L.mas_set_tuning2(fused, 0)
t_ys, t_xs = make_lengths(np.random.default_rng(0), B, T_y, T_x, False)
ty, tx = torch.as_tensor(t_ys).cuda(), torch.as_tensor(t_xs).cuda()
g = torch.Generator(device='cuda').manual_seed(1)
bufs = [torch.randn(B, T_y, T_x, generator=g, device='cuda') * 20 - 400 for _ in range(3)]
for i in range(2): vits_b200.maximum_path_from_lengths(bufs[i], ty, tx)
torch.cuda.synchronize()
tr = torch.zeros(8 * 512 * 2, dtype=torch.int64, device='cuda')
L.mas_set_trace(tr.data_ptr())
vits_b200.maximum_path_from_lengths(bufs[2], ty, tx)
torch.cuda.synchronize()
L.mas_set_trace(None)
t = tr.cpu().numpy().astype(np.uint64).reshape(8, 512, 2)
names = {1: 'stage top', 2: 'got full', 3: 'got bfull', 4: 'rows done', 5: 'flushed', 6: 'released'}
t0 = min(int(t[w, 0, 1]) for w in range(8) if t[w, 0, 1])
for w in range(4):
    if not t[w, 0, 1]: continue
    ev = [(int(x[0]) >> 32, int(x[0]) & 0xffffffff, int(x[1]) - t0) for x in t[w] if x[1]]
    print(f"warp {w}: {len(ev)} events, last at {ev[-1][2]}")
    # per-stage breakdown for stages 10..13
    for tag, idx, tt in ev:
        if 10 <= idx <= 12: print(f"   {tt:8d} {names[tag]} {idx}")
    tops = [tt for tag, idx, tt in ev if tag == 1]
    d = np.diff(tops)
    print("   stage periods:", d.tolist())
