"""Kernel-level timing sweep (GPU box): per-kernel durations for tuning combinations.
usage: python tools/sweep.py [c2|c3|c4] [--ragged]"""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vits_b200
from vits_b200 import _lib
from bench import WORKLOADS, make_lengths

def time_call(fn, bufs, reps=5):
    """GPU-bound timing: the calls are captured into one CUDA graph (no Python/ctypes cost per call)."""
    for i in range(2): fn(bufs[i % len(bufs)])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for nc in bufs: fn(nc)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * len(bufs)) * 1e3  # us

def main():
    wl = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith('-') else 'c2'
    ragged = '--ragged' in sys.argv
    B, T_y, T_x = WORKLOADS[wl]
    L = _lib.lib()
    rng = np.random.default_rng(1234)
    t_ys, t_xs = make_lengths(rng, B, T_y, T_x, ragged)
    ty, tx = torch.as_tensor(t_ys).cuda(), torch.as_tensor(t_xs).cuda()
    g = torch.Generator(device='cuda').manual_seed(1)
    bufs = [torch.randn(B, T_y, T_x, generator=g, device='cuda') * 20 - 400 for _ in range(6)]
    full = lambda nc: vits_b200.maximum_path_from_lengths(nc, ty, tx)
    print(f"workload {wl} B={B} T_y={T_y} T_x={T_x} ragged={ragged}")
    Ks = [int(a[2:]) for a in sys.argv if a.startswith('-K')] or [1, 2, 4, 8]
    Rs = [int(a[2:]) for a in sys.argv if a.startswith('-R')] or [8, 16, 32]
    Fs = [int(a[2:]) for a in sys.argv if a.startswith('-F')] or [0, 1]
    Hs = [int(a[2:]) for a in sys.argv if a.startswith('-H')] or [0]
    for K, R, S, F, H in itertools.product(Ks, Rs, (0,), Fs, Hs):
        L.mas_set_tuning(K, R, S, 0)
        L.mas_set_tuning2(F, H)
        try:
            L.mas_set_debug_kernels(1); t1 = time_call(full, bufs)
            L.mas_set_debug_kernels(3); t3 = time_call(full, bufs)
            L.mas_set_debug_kernels(7); t7 = time_call(full, bufs)
            L.mas_set_tuning(K, R, S, 1); t7p = time_call(full, bufs)
            print(f"K={K} R={R:2d} fused={F} H={H}: forward {t1:7.1f} us | +backtrack {t3:7.1f} | +writeout {t7:7.1f} us | all PDL {t7p:7.1f} us")
        except Exception as e:
            print(f"K={K} R={R} failed: {e}")
    L.mas_set_debug_kernels(7); L.mas_set_tuning(0, 0, 0, -1); L.mas_set_tuning2(-1, 0)

main()
