"""Kernel spans (globaltimer) of one call per backtrack mode.  usage: tl_modes.py [c2] [mask|lens]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, vits_b200
from vits_b200 import _lib
from bench import WORKLOADS, make_lengths
L = _lib.lib()
wl = sys.argv[1] if len(sys.argv) > 1 else 'c2'
use_mask = len(sys.argv) > 2 and sys.argv[2] == 'mask'
index_only = len(sys.argv) > 3 and sys.argv[3] == 'index'
B, T_y, T_x = WORKLOADS[wl]
t_ys, t_xs = make_lengths(np.random.default_rng(0), B, T_y, T_x, False)
ty, tx = torch.as_tensor(t_ys).cuda(), torch.as_tensor(t_xs).cuda()
mask = ((torch.arange(T_y, device='cuda')[None, :] < ty[:, None])[:, :, None] & (torch.arange(T_x, device='cuda')[None, :] < tx[:, None])[:, None, :]).float()
g = torch.Generator(device='cuda').manual_seed(1)
bufs = [torch.randn(B, T_y, T_x, generator=g, device='cuda') * 20 - 400 for _ in range(3)]
tl = torch.zeros(16, dtype=torch.int64, device='cuda')
names = ["dp_start", "dp_done", "fwd_end", "bt_start", "bt_end", "fill_start", "fill_done", "wo_end"]
call = (lambda nc: vits_b200.maximum_path(nc, mask)) if use_mask else (lambda nc: vits_b200.maximum_path_from_lengths(nc, ty, tx))
if index_only: call = lambda nc: vits_b200.maximum_path_index(nc, y_lengths=ty, x_lengths=tx)
for mode in (1, 2, 3):
    L.mas_set_tuning2(mode, 0)
    try:
        for i in range(2): call(bufs[i])
    except Exception as ex:
        print(mode, "unsupported"); continue
    res = []
    for rep in range(5):
        tl.zero_(); tl[0] = tl[3] = tl[5] = -1; torch.cuda.synchronize()
        L.mas_set_timeline(tl.data_ptr()); call(bufs[rep % 3]); torch.cuda.synchronize(); L.mas_set_timeline(None)
        v = tl.cpu().numpy().astype(np.uint64); t0 = int(v[0])
        res.append([(int(x) - t0) / 1e3 if int(x) not in (0, 2**64 - 1) else float('nan') for x in v])
    import warnings; warnings.simplefilter("ignore")
    r = np.nanmedian(np.array(res), 0)
    print(f"mode={mode} {'mask' if use_mask else 'lens'}: " + "  ".join(f"{n}={x:.1f}" for n, x in zip(names, r)))
L.mas_set_tuning2(-1, 0)
