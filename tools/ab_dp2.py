"""A/B of the wavefront DP kernel generations (mas_set_tuning3 `wavefront`: 1 = first generation, 16/17/32/33 =
32 = mas_dp2, 33 = mas_dp2 with the instruction-cache warmer; ring mode 1/2 = skew): bit-exactness against the CPU
oracle on a shape sweep, then the period of back-to-back calls (one graph of 8 calls over 4 rotating buffers).
usage: python tools/ab_dp2.py [c2|c3] [--no-fuzz]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import vits_b200
import vits_b200.monotonic_align as mp
from vits_b200 import _lib
from bench import WORKLOADS, make_lengths
from oracle import mas_oracle
from helpers import random_lengths, path_to_index

MODES = [(1, 0), (32, 0), (33, 0), (33, 1), (1, 1)]   # (wavefront mode, ring mode = skew; 0 = automatic)
if "--modes" in sys.argv:
    MODES = [tuple(int(y) for y in x.split(":")) for x in sys.argv[sys.argv.index("--modes") + 1].split(",")]
L = _lib.lib()
if "--pdl1" in sys.argv:   # the forward kernel launched ordinarily (mas_set_tuning pdl = 1; default: behind the previous call)
    L.mas_set_tuning(0, 0, 0, 1)
if "--no-fuzz" not in sys.argv:
    rng = np.random.default_rng(5)
    shapes = [(2, 900, tx) for tx in (24, 64, 100, 128, 192, 256, 300, 384, 512)] + [(3, 333, 300), (2, 64, 60), (2, 33, 20), (1, 5, 4),
              (2, 2000, 200), (70, 500, 100), (5, 1024, 192)]
    for shape in shapes:
        B, T_y, T_x = shape
        res = []
        for kind in ("normal", "subsentinel", "ties"):
            if kind == "normal": nc = (rng.standard_normal(shape) * 3 - 4).astype(np.float32)
            elif kind == "subsentinel": nc = (rng.standard_normal(shape) * 1e8 - 2e9).astype(np.float32)
            else: nc = rng.integers(-3, 1, size=shape).astype(np.float32)
            t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
            if B > 1: t_ys[0], t_xs[0] = T_y, min(T_x, T_y)
            want = path_to_index(mas_oracle.maximum_path_numpy(nc, t_ys, t_xs))
            ncd = torch.from_numpy(nc).cuda()
            for wf, ring in MODES:
                L.mas_set_tuning3(wf, ring, 0, 0)
                nbad = 0
                for rep in range(3):
                    got = mp.maximum_path_index(ncd, y_lengths=torch.as_tensor(t_ys), x_lengths=torch.as_tensor(t_xs)).cpu().numpy()
                    nbad += int((got != want).any())
                res.append(f"{wf}/{ring}:{nbad}")
        print(shape, "bad reps (normal | sub-sentinel | ties):", " ".join(res), flush=True)
    L.mas_set_tuning3(-1, 0, 0, 0)

for wl in [a for a in sys.argv[1:] if a in WORKLOADS] or ["c2"]:
    B, T_y, T_x = WORKLOADS[wl]
    g = torch.Generator(device='cuda').manual_seed(1)
    bufs = [torch.randn(B, T_y, T_x, generator=g, device='cuda') * 20 - 400 for _ in range(4)]
    for ragged in (True, False):
        t_ys, t_xs = make_lengths(np.random.default_rng(1234), B, T_y, T_x, ragged)
        ty, tx = torch.as_tensor(t_ys).cuda(), torch.as_tensor(t_xs).cuda()
        mask = ((torch.arange(T_y, device='cuda')[None, :] < ty[:, None])[:, :, None]
                & (torch.arange(T_x, device='cuda')[None, :] < tx[:, None])[:, None, :]).float()
        ref = None
        for wf, ring in MODES * 2:
            L.mas_set_tuning3(wf, ring, 0, 0)
            outs = [vits_b200.maximum_path_index(bufs[i], y_lengths=ty, x_lengths=tx) for i in range(2)]
            torch.cuda.synchronize()
            if ref is None: ref = [o.clone() for o in outs]
            ok = all(torch.equal(a, b) for a, b in zip(outs, ref))
            res = []
            for use_mask in (False, True):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    for i in range(8):
                        o = vits_b200.maximum_path(bufs[i % 4], mask) if use_mask else vits_b200.maximum_path_from_lengths(bufs[i % 4], ty, tx)
                for _ in range(3): gr.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(25): gr.replay()
                e1.record(); torch.cuda.synchronize()
                res.append(e0.elapsed_time(e1) * 1e3 / 200)
            print(f"{wl} {'ragged' if ragged else 'full  '} wf={wf:2d} skew={ring}: lens {res[0]:7.2f} us/call  mask {res[1]:7.2f} us/call  same index as first: {ok}", flush=True)
    L.mas_set_tuning3(-1, 0, 0, 0)
