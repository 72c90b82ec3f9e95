"""Stress of the streamed stats->path at the SMALL shapes of tests/test_neg_cent_gpu.py (test_streamed_stats_to_path_is_bit_identical),
exactly as the test calls it (two streamed calls back to back, dense then index), many repetitions, reporting what differs.
usage: stress_streamed_small.py [reps]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import vits_b200
from helpers import random_lengths, path_to_index
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
shapes = [(2, 192, 130, 40), (3, 192, 260, 90), (4, 80, 300, 64), (5, 192, 700, 192), (3, 192, 520, 300), (2, 192, 1100, 256)]
def inputs(B, C, T_y, T_x, seed, t_ys, t_xs):
    g = torch.Generator(device="cuda").manual_seed(seed)
    z = torch.randn(B, C, T_y, generator=g, device="cuda"); m = torch.randn(B, C, T_x, generator=g, device="cuda")
    ls = torch.randn(B, C, T_x, generator=g, device="cuda") * 0.3
    for b in range(B):
        z[b, :, int(t_ys[b]):] = 0; m[b, :, int(t_xs[b]):] = 0; ls[b, :, int(t_xs[b]):] = 0
    return z, m, ls
for ragged in (True, False):
    for shape in shapes:
        B, C, T_y, T_x = shape
        rng = np.random.default_rng(sum(shape))
        t_ys, t_xs = random_lengths(rng, B, T_y, T_x) if ragged else (np.full(B, T_y, np.int32), np.full(B, T_x, np.int32))
        z, m, ls = inputs(B, C, T_y, T_x, 3, t_ys, t_xs)
        ty, tx = torch.as_tensor(t_ys).cuda(), torch.as_tensor(t_xs).cuda()
        want = vits_b200.maximum_path_from_stats(z, m, ls, tx, ty, index=True, streamed=False)
        torch.cuda.synchronize()
        bad = 0
        for rep in range(reps):
            got = vits_b200.maximum_path_from_stats(z, m, ls, tx, ty, streamed=True)
            idx = vits_b200.maximum_path_from_stats(z, m, ls, tx, ty, index=True, streamed=True)
            if rep % 4 == 3:
                torch.cuda.synchronize()
            st = vits_b200.status_nosync()
            d1 = (idx != want).any(1).nonzero().flatten().tolist()
            dsum = int(got.sum())
            gi = got.argmax(2).int(); gi[got.sum(2) == 0] = -1
            d2 = (gi != want).any(1).nonzero().flatten().tolist()
            if d1 or d2 or st or dsum != int(t_ys.sum()):
                bad += 1
                if bad <= 5:
                    b0 = (d1 + d2)[0] if (d1 + d2) else -1
                    fr = ((idx[b0] != want[b0]) | (gi[b0] != want[b0])).nonzero().flatten() if b0 >= 0 else []
                    print(f"  {shape} ragged={ragged} rep {rep}: index differs in {d1}, dense in {d2}, dense sum {dsum}/{int(t_ys.sum())}, status {st}; "
                          f"b={b0} t_y={t_ys[b0]} t_x={t_xs[b0]} frames {len(fr)} [{int(fr[0]) if len(fr) else -1}..{int(fr[-1]) if len(fr) else -1}]", flush=True)
                vits_b200.status_nosync(reset=True)
        print(f"{shape} ragged={ragged}: {bad} of {reps} repetitions differ", flush=True)
