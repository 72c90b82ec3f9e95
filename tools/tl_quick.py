import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, vits_b200
from vits_b200 import _lib
from bench import WORKLOADS, make_lengths
L = _lib.lib()
B, T_y, T_x = WORKLOADS['c2']
t_ys, t_xs = make_lengths(np.random.default_rng(0), B, T_y, T_x, False)
ty, tx = torch.as_tensor(t_ys).cuda(), torch.as_tensor(t_xs).cuda()
g = torch.Generator(device='cuda').manual_seed(1)
bufs = [torch.randn(B, T_y, T_x, generator=g, device='cuda') * 20 - 400 for _ in range(3)]
tl = torch.zeros(8, dtype=torch.int64, device='cuda')
for H in (1, 2, 3, 4):
    L.mas_set_tuning2(1, H); L.mas_set_tuning(0, 0, 0, 0)
    for i in range(2): vits_b200.maximum_path_from_lengths(bufs[i], ty, tx)
    res = []
    for rep in range(3):
        tl.zero_(); tl[0] = tl[5] = -1; torch.cuda.synchronize()
        L.mas_set_timeline(tl.data_ptr()); vits_b200.maximum_path_from_lengths(bufs[2], ty, tx); torch.cuda.synchronize(); L.mas_set_timeline(None)
        v = tl.cpu().numpy().astype(np.uint64); t0 = int(v[0])
        res.append([(int(x) - t0) / 1e3 for x in v[1:5]])
    r = np.median(np.array(res), 0)
    print(f"H={H}: dp_done {r[0]:.1f}  helpers_done {r[2]:.1f}  chain_done {r[3]:.1f}  fwd_end {r[1]:.1f} us")
L.mas_set_tuning2(-1, 0); L.mas_set_tuning(0, 0, 0, 1)
