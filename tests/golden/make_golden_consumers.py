"""Golden vectors for commons.generate_path from the REAL reference (run in the build container only):
    python tests/golden/make_golden_consumers.py     -> tests/golden/consumers_golden.npz
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
import commons  # noqa: E402  (the reference's own module)

rng = np.random.default_rng(2024)
out = {}
for name, (b, t_x) in {"small": (2, 7), "medium": (3, 40)}.items():
    dur = np.ceil(np.exp(rng.standard_normal((b, 1, t_x))) * 1.5).astype(np.float32)
    x_len = rng.integers(t_x // 2 + 1, t_x + 1, size=b)
    x_len[0] = t_x
    x_mask = (np.arange(t_x)[None, :] < x_len[:, None]).astype(np.float32)[:, None, :]
    dur = dur * x_mask
    y_len = np.maximum(dur.sum((1, 2)), 1).astype(np.int64)
    t_y = int(y_len.max())
    y_mask = (np.arange(t_y)[None, :] < y_len[:, None]).astype(np.float32)[:, None, :]
    attn_mask = torch.from_numpy(x_mask).unsqueeze(2) * torch.from_numpy(y_mask).unsqueeze(-1)
    path = commons.generate_path(torch.from_numpy(dur), attn_mask)
    out[f"{name}/duration"] = dur
    out[f"{name}/mask"] = attn_mask.numpy()
    out[f"{name}/path"] = path.numpy()
out["names"] = np.array(["small", "medium"])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "consumers_golden.npz"), **out)
print("wrote consumers_golden.npz", {k: v.shape for k, v in out.items() if k != "names"})
