"""Per-superstep cycle stamps of the wavefront DP kernel's CTA 0 (needs a -DMAS_TRACE build: tools/build_variant.py
trace -DMAS_TRACE; run with VITS_MAS_LIB=vits_b200/build_trace/libvits_mas_trace.so).
usage: trace_dp.py [c2] [ragged 0/1] [alone 0/1] [wavefront mode of mas_set_tuning3]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vits_b200
from vits_b200 import _lib
from bench import WORKLOADS, make_lengths
wl = sys.argv[1] if len(sys.argv) > 1 else 'c2'
ragged = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
alone = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
wf = int(sys.argv[4]) if len(sys.argv) > 4 else -1
use_mask = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
B, T_y, T_x = WORKLOADS[wl]
L = _lib.lib()
L.mas_set_tuning3(wf, 0, 0, 0)
t_ys, t_xs = make_lengths(np.random.default_rng(0), B, T_y, T_x, ragged)
ty, tx = torch.as_tensor(t_ys).cuda(), torch.as_tensor(t_xs).cuda()
g = torch.Generator(device='cuda').manual_seed(1)
bufs = [torch.randn(B, T_y, T_x, generator=g, device='cuda') * 20 - 400 for _ in range(3)]
if use_mask:
    mask = ((torch.arange(T_y, device='cuda')[None, :] < ty[:, None])[:, :, None]
            & (torch.arange(T_x, device='cuda')[None, :] < tx[:, None])[:, None, :]).float()
def call(i):
    return vits_b200.maximum_path(bufs[i], mask) if use_mask else vits_b200.maximum_path_from_lengths(bufs[i], ty, tx)
if alone: L.mas_set_debug_kernels(1)   # forward kernel only
for i in range(2): call(i)
torch.cuda.synchronize()
n = 8 * 256 * 8 + 2 * B
tr = torch.zeros(n, dtype=torch.int64, device='cuda')
L.mas_set_trace(tr.data_ptr())
call(2)
torch.cuda.synchronize()
L.mas_set_trace(None)
L.mas_set_debug_kernels(7)
t = tr.cpu().numpy()
gt = t[8 * 256 * 8:].reshape(B, 2)
print(f"{wl} ragged={ragged} alone={alone} wf={wf} mask={use_mask}: per-CTA DP duration (us): b0 {(gt[0,1]-gt[0,0])/1e3:.1f}  max {((gt[:,1]-gt[:,0]).max())/1e3:.1f}  "
      f"first start -> last end {(gt[:,1].max()-gt[:,0].min())/1e3:.1f}")
t = t[:8 * 256 * 8].reshape(8, 256, 8)
for w in range(8):
    rows = [(s, t[w, s]) for s in range(256) if t[w, s, 0]]
    if not rows: continue
    start = np.array([r[1][0] for r in rows]); body = np.array([r[1][3] for r in rows]); end = np.array([r[1][7] for r in rows])
    blocked = np.array([int(r[1][1]) - int(r[1][2]) if r[1][1] else 0 for r in rows])   # time inside a blocking start BEFORE the superstep
    per = np.diff(start)
    print(f"warp {w}: {len(rows)} supersteps, first start -> last end {end[-1]-start[0]} cycles; medians: period {np.median(per):.0f}  "
          f"block {np.median(body-start):.0f}  tail {np.median(end-body):.0f}  gap-to-next {np.median(start[1:]-end[:-1]):.0f}; blocked in total {blocked.sum()}")
    print("   period :", per.tolist())
    print("   block  :", (body - start).tolist())
    print("   blocked:", blocked.tolist())
    fx = [(r[0], int(r[1][5]) - int(r[1][4]), int(r[1][6]) - int(r[1][5])) for r in rows if r[1][4]]
    if fx: print("   diagonal fix-up (superstep, cycles waiting for its chunks, cycles zeroing + fence + reload):", fx)
