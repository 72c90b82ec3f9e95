"""Randomised bit-exactness sweep over shapes / ring modes (debug)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import vits_b200.monotonic_align as mp
from oracle import mas_oracle
from helpers import random_lengths, path_to_index
L = mp._lib.lib()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
shapes = [(2, 900, tx) for tx in (64, 128, 192, 256, 257, 300, 320, 384, 449, 512, 777)] + [(3, 333, 300), (2, 64, 60), (2, 2000, 200), (70, 500, 100)]
for shape in shapes:
    B, T_y, T_x = shape
    nc = (rng.standard_normal(shape) * 3 - 4).astype(np.float32)
    t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
    want = path_to_index(mas_oracle.maximum_path_numpy(nc, t_ys, t_xs))
    ncd = torch.from_numpy(nc).cuda()
    res = []
    for mode in (0,):
        for stream in (0, 1, 2, 3, -1):
            L.mas_set_tuning(0, 0, 0, -1); L.mas_set_tuning2(stream, 0)
            nbad = 0; err = ""
            for rep in range(4):
                try:
                    got = mp.maximum_path_index(ncd, y_lengths=torch.as_tensor(t_ys), x_lengths=torch.as_tensor(t_xs)).cpu().numpy()
                    nbad += int((got != want).any())
                except Exception as ex:
                    err = "E"; break
            res.append(err or str(nbad))
    print(shape, "lens", t_ys.tolist(), t_xs.tolist(), "| bad reps per (mode,stream):", " ".join(res))
