"""Small run of every maximum_path mode for compute-sanitizer (memcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import vits_b200, vits_b200.monotonic_align as mp
from oracle import mas_oracle
from helpers import random_lengths, path_to_index
L = mp._lib.lib()
rng = np.random.default_rng(3)
for shape in [(3, 150, 96), (2, 260, 129), (2, 200, 31), (2, 90, 300)]:
    B, T_y, T_x = shape
    nc = (rng.standard_normal(shape) * 2).astype(np.float32)
    t_ys, t_xs = random_lengths(rng, B, T_y, min(T_x, T_y))
    t_xs = np.minimum(t_xs, t_ys).astype(np.int32)
    want = mas_oracle.maximum_path_numpy(nc, t_ys, t_xs)
    ncd = torch.from_numpy(nc).cuda()
    for mode in (-1, 0, 1, 2, 3):
        L.mas_set_tuning2(mode, 0)
        got = mp.maximum_path_from_lengths(ncd, torch.as_tensor(t_ys), torch.as_tensor(t_xs))
        idx = mp.maximum_path_index(ncd, y_lengths=torch.as_tensor(t_ys), x_lengths=torch.as_tensor(t_xs))
        torch.cuda.synchronize()
        assert np.array_equal(got.cpu().numpy().astype(np.int32), want), (shape, mode)
        assert np.array_equal(idx.cpu().numpy(), path_to_index(want)), (shape, mode)
L.mas_set_tuning2(-1, 0)
idx = idx.cuda() if not idx.is_cuda else idx
w = vits_b200.path_durations(idx, T_x)
m, l = vits_b200.expand_prior(idx, torch.randn(B, 7, T_x, device="cuda"), torch.randn(B, 7, T_x, device="cuda"))
z = vits_b200.neg_cent(torch.randn(2, 192, 100, device="cuda"), torch.randn(2, 192, 40, device="cuda"), torch.randn(2, 192, 40, device="cuda") * 0.3)
torch.cuda.synchronize()
print("sanitize run ok")
