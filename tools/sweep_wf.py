"""Period of back-to-back maximum_path calls (one graph of 8 calls over 4 rotating buffers, 25 replays) for the
wavefront kernel's ring modes / columns per lane / ring depths.  usage: python tools/sweep_wf.py [c2|c3|c4]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vits_b200
from vits_b200 import _lib
from bench import WORKLOADS, make_lengths

wl = next((a for a in sys.argv[1:] if a in WORKLOADS), 'c2')
B, T_y, T_x = WORKLOADS[wl]
L = _lib.lib()
g = torch.Generator(device='cuda').manual_seed(1)
bufs = [torch.randn(B, T_y, T_x, generator=g, device='cuda') * 20 - 400 for _ in range(4)]
ref = {}
for ragged in (True, False):
    t_ys, t_xs = make_lengths(np.random.default_rng(1234), B, T_y, T_x, ragged)
    ty, tx = torch.as_tensor(t_ys).cuda(), torch.as_tensor(t_xs).cuda()
    cfgs = [(0, 0, 0), (1, 2, 0), (2, 2, 0), (1, 2, 4), (1, 2, 5), (1, 1, 0), (2, 1, 0), (1, 4, 0), (4, 2, 0), (4, 1, 0)]
    if '--ab' in sys.argv:   # default vs skew 2 (and 3), alternating
        cfgs = [(1, 0, 0), (2, 0, 0), (3, 0, 0)] * 4
    if '--mask' in sys.argv:   # lengths given vs lengths taken from a [B,T_y,T_x] fp32 mask, default configuration
        cfgs = [(0, 0, 0), (0, 0, -1)] * 4
        mask = ((torch.arange(T_y, device='cuda')[None, :] < ty[:, None])[:, :, None]
                & (torch.arange(T_x, device='cuda')[None, :] < tx[:, None])[:, None, :]).float()
    for ring, K, S in cfgs:
        use_mask = S == -1
        S = max(S, 0)
        L.mas_set_tuning3(-1, ring, S, K)
        try:
            outs = [vits_b200.maximum_path_index(bufs[i], y_lengths=ty, x_lengths=tx) for i in range(2)]
            torch.cuda.synchronize()
            key = ragged
            if key not in ref: ref[key] = [o.clone() for o in outs]
            ok = all(torch.equal(a, b) for a, b in zip(outs, ref[key]))
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for i in range(8):
                    if use_mask:
                        o = vits_b200.maximum_path(bufs[i % 4], mask)
                    else:
                        o = vits_b200.maximum_path_from_lengths(bufs[i % 4], ty, tx)
            for _ in range(3): gr.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(25): gr.replay()
            e1.record(); torch.cuda.synchronize()
            print(f"{wl} {'ragged' if ragged else 'full  '} ring={ring} K={K} S={S} {'mask' if use_mask else 'lens'}: {e0.elapsed_time(e1) * 1e3 / 200:7.2f} us/call  same index as default: {ok}")
        except Exception as ex:
            print(f"{wl} ring={ring} K={K} S={S}: {type(ex).__name__} {ex}")
L.mas_set_tuning3(-1, 0, 0, 0)
