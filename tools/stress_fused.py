"""Repeat the streamed stats->path at the shape that once differed (c3 ragged, profiles/r02q_streamed_vs_two_call.txt) and report
every repetition whose path is not bit-identical to the two-call form."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vits_b200
from bench import make_lengths
from check_fused import inputs
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
for (B, C, T_y, T_x) in [(32, 192, 1536, 256), (64, 192, 1024, 192), (8, 192, 700, 192)]:
    t_ys, t_xs = make_lengths(np.random.default_rng(B + T_y), B, T_y, T_x, True)
    z, m, ls = inputs(B, C, T_y, T_x, 7, t_ys, t_xs)
    ty, tx = torch.as_tensor(t_ys).cuda(), torch.as_tensor(t_xs).cuda()
    want = vits_b200.maximum_path_from_stats(z, m, ls, tx, ty, index=True, streamed=False)
    torch.cuda.synchronize()
    bad = 0
    for rep in range(reps):
        if rep % 3 == 2:   # vary what runs before the call
            junk = torch.randn(1 << (20 + rep % 5), device='cuda').sum()
        got = vits_b200.maximum_path_from_stats(z, m, ls, tx, ty, index=True, streamed=True)
        dense = vits_b200.maximum_path_from_stats(z, m, ls, tx, ty, streamed=True) if rep % 2 else None
        torch.cuda.synchronize()
        diff = (got != want).any(1).nonzero().flatten().tolist()
        dsum = int(dense.sum()) if dense is not None else int(t_ys.sum())
        if diff or dsum != int(t_ys.sum()):
            bad += 1
            b0 = diff[0] if diff else -1
            d = (got[b0] != want[b0]).nonzero().flatten() if diff else []
            print(f"  {(B, C, T_y, T_x)} rep {rep}: utterances {diff[:6]} differ; first b={b0} t_y={t_ys[b0]} t_x={t_xs[b0]} frames {len(d)} "
                  f"[{int(d[0]) if len(d) else -1}..{int(d[-1]) if len(d) else -1}] dense sum {dsum}/{int(t_ys.sum())} status {vits_b200.status_nosync()}", flush=True)
    print(f"{(B, C, T_y, T_x)}: {bad} of {reps} repetitions differ", flush=True)
