"""Where the time between two back-to-back maximum_path calls goes: two calls captured in ONE graph (so the
programmatic edge between them is kept), each stamping its own %globaltimer timeline (mas_set_timeline).
usage: python tools/timeline_gap.py [c2|c3|c4] [--ragged] [--pdl0 | --xpdl] [--fence]
--xpdl: forward kernel launched programmatically behind the previous call (the former default)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vits_b200
from vits_b200 import _lib
from bench import WORKLOADS, make_lengths

wl = next((a for a in sys.argv[1:] if a in WORKLOADS), 'c2')
B, T_y, T_x = WORKLOADS[wl]
L = _lib.lib()
t_ys, t_xs = make_lengths(np.random.default_rng(0), B, T_y, T_x, '--ragged' in sys.argv)
ty, tx = torch.as_tensor(t_ys).cuda(), torch.as_tensor(t_xs).cuda()
g = torch.Generator(device='cuda').manual_seed(1)
bufs = [torch.randn(B, T_y, T_x, generator=g, device='cuda') * 20 - 400 for _ in range(4)]
names = ['fwd first start (after griddepcontrol.wait)', 'fwd last DP warp done', 'fwd last end', 'bt first start',
         'bt last end', 'wo first start', 'wo last zero-fill done', 'lengths known']
N = 4
tls = [torch.zeros(16, dtype=torch.int64, device='cuda') for _ in range(N)]
def reset():
    for tl in tls:
        tl.zero_(); tl[0] = tl[3] = tl[5] = -1
L.mas_set_tuning(0, 0, 0, 0 if '--pdl0' in sys.argv else 2 if '--xpdl' in sys.argv else 1 if '--pdl1' in sys.argv else -1)
use_mask = '--mask' in sys.argv   # lengths from a [B,T_y,T_x] fp32 mask (the reference's signature) instead of given
if use_mask:
    mask = ((torch.arange(T_y, device='cuda')[None, :] < ty[:, None])[:, :, None]
            & (torch.arange(T_x, device='cuda')[None, :] < tx[:, None])[:, None, :]).float()
def call(i):
    return vits_b200.maximum_path(bufs[i], mask) if use_mask else vits_b200.maximum_path_from_lengths(bufs[i], ty, tx)
for i in range(2): call(i)
torch.cuda.synchronize()
dummy = torch.zeros(1, device='cuda')
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for i in range(N):
        L.mas_set_timeline(tls[i].data_ptr())
        call(i)
        if '--fence' in sys.argv:   # an ordinary launch between the calls: no programmatic edge, no early residency
            dummy.add_(1)
L.mas_set_timeline(None)
gr.replay(); torch.cuda.synchronize()
for rep in range(2):
    reset(); gr.replay(); torch.cuda.synchronize()
    v = [tl.cpu().numpy().astype(np.uint64) for tl in tls]
    t0 = int(v[0][0])
    print(f"{wl} rep {rep}: {N} calls in one graph, us relative to call 0's forward start")
    for i in range(N):
        row = "  ".join(f"{n.split(' (')[0]}={(int(x) - t0) / 1e3:7.2f}" for n, x in zip(names, v[i]) if int(x) not in (0, 2**64 - 1))
        print(f"  call {i}: {row}")
        if int(v[i][8]):
            print(f"          after the last DP warp: top group's words seen +{(int(v[i][8]) - int(v[i][1])) / 1e3:5.2f}, all tables done +{(int(v[i][9]) - int(v[i][1])) / 1e3:5.2f}, "
                  f"last tabulated group's words complete +{(int(v[i][13]) - int(v[i][1])) / 1e3:5.2f}, its table +{(int(v[i][11]) - int(v[i][1])) / 1e3:5.2f}, top groups walked +{(int(v[i][12]) - int(v[i][1])) / 1e3:5.2f}, chain over groups done +{(int(v[i][10]) - int(v[i][1])) / 1e3:5.2f}, last backtrack CTA past its index/ones +{(int(v[i][4]) - int(v[i][1])) / 1e3:5.2f} us")
        print(f"          DP {(int(v[i][1]) - int(v[i][0])) / 1e3:6.2f} us, backtrack tail {(int(v[i][4]) - int(v[i][1])) / 1e3:6.2f} us, fill {(int(v[i][6]) - int(v[i][5])) / 1e3:6.2f} us")
    for i in range(1, N):
        last_end = max(int(v[i - 1][4]), int(v[i - 1][2]))
        print(f"  call {i-1} -> {i}: period {(int(v[i][0]) - int(v[i-1][0])) / 1e3:6.2f} us, "
              f"previous chain's last stamp -> forward past its wait {(int(v[i][0]) - last_end) / 1e3:6.2f} us")
