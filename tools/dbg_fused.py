"""Where the streamed stats->path differs from the two-call path (c3 ragged showed 49 missing ones in r02q)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vits_b200
from bench import make_lengths
from check_fused import inputs
B, C, T_y, T_x = (32, 192, 1536, 256) if len(sys.argv) < 5 else tuple(int(a) for a in sys.argv[1:5])
t_ys, t_xs = make_lengths(np.random.default_rng(B + T_y), B, T_y, T_x, True)
z, m, ls = inputs(B, C, T_y, T_x, 7, t_ys, t_xs)
ty, tx = torch.as_tensor(t_ys).cuda(), torch.as_tensor(t_xs).cuda()
want = vits_b200.maximum_path_from_stats(z, m, ls, tx, ty, index=True, streamed=False)
torch.cuda.synchronize()
for rep in range(4):
    got = vits_b200.maximum_path_from_stats(z, m, ls, tx, ty, index=True, streamed=True)
    dense = vits_b200.maximum_path_from_stats(z, m, ls, tx, ty, streamed=True)
    torch.cuda.synchronize()
    bad = (got != want).any(1).nonzero().flatten().tolist()
    print(f"rep {rep}: utterances with a differing index: {bad}; dense sum {int(dense.sum())} want {int(t_ys.sum())}; status {vits_b200.status_nosync()}")
    for b in bad[:4]:
        d = (got[b] != want[b]).nonzero().flatten()
        print(f"   b={b} t_y={t_ys[b]} t_x={t_xs[b]} differing frames {len(d)}: first {int(d[0])} last {int(d[-1])}; got {got[b, d[:6]].tolist()} want {want[b, d[:6]].tolist()}")
    dsum = dense.sum((1, 2)).cpu().numpy().astype(int)
    wrong = [(b, int(dsum[b]), int(t_ys[b])) for b in range(B) if dsum[b] != t_ys[b]]
    print("   dense row sums off:", wrong[:8])
