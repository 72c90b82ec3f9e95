"""Cycles per superstep of the wavefront DP kernels in production builds (no stamps): the slope of the call time
over T_y at fixed B = 64, T_x = 192, full lengths given (one graph of 8 calls over 4 rotating buffers)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vits_b200
from vits_b200 import _lib
L = _lib.lib()
B, T_x = 64, 192
res = {}
MODES = ((33, 1),) if '--dp2' in sys.argv else ((1, 0), (33, 1), (33, 2), (32, 1))
for T_y in (256, 512, 1024):
    g = torch.Generator(device='cuda').manual_seed(1)
    bufs = [torch.randn(B, T_y, T_x, generator=g, device='cuda') * 20 - 400 for _ in range(4)]
    ty = torch.full((B,), T_y, dtype=torch.int32, device='cuda'); tx = torch.full((B,), T_x, dtype=torch.int32, device='cuda')
    for wf, ring in MODES:
        L.mas_set_tuning3(wf, ring, 0, 0)
        for i in range(2): vits_b200.maximum_path_from_lengths(bufs[i], ty, tx)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for i in range(8): vits_b200.maximum_path_from_lengths(bufs[i % 4], ty, tx)
        for _ in range(3): gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(25): gr.replay()
        e1.record(); torch.cuda.synchronize()
        res[(wf, ring, T_y)] = e0.elapsed_time(e1) * 1e3 / 200
        print(f"T_y={T_y} wf={wf} skew={ring}: {res[(wf, ring, T_y)]:7.2f} us/call", flush=True)
L.mas_set_tuning3(-1, 0, 0, 0)
for wf, ring in MODES:
    for a, b_ in ((256, 512), (512, 1024)):
        d = res[(wf, ring, b_)] - res[(wf, ring, a)]
        print(f"wf={wf} skew={ring}: {a}->{b_}: +{d:6.2f} us = {d * 1e3 / ((b_ - a) / 32):6.1f} ns per superstep (~{d * 1965 / ((b_ - a) / 32):5.0f} cycles at 1965 MHz)")
