"""Quick numerical check of the tcgen05 neg_cent against the fp32 torch expression (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import vits_b200
from oracle import mas_oracle
L = vits_b200._lib.lib()
shapes = [(1, 32, 128, 16), (1, 192, 128, 192), (2, 192, 300, 100), (2, 192, 1024, 192), (1, 192, 777, 257), (1, 80, 130, 70)]
if len(sys.argv) > 1: shapes = shapes[:int(sys.argv[1])]
for (B, C, T_y, T_x) in shapes:
    g = torch.Generator(device='cuda').manual_seed(B + C + T_y + T_x)
    z = torch.randn(B, C, T_y, generator=g, device='cuda')
    m = torch.randn(B, C, T_x, generator=g, device='cuda')
    ls = torch.randn(B, C, T_x, generator=g, device='cuda') * 0.3
    ref = mas_oracle.neg_cent_torch(z, m, ls)
    for impl in (1, 0):
        L.mas_set_neg_cent_impl(impl)
        out = vits_b200.neg_cent(z, m, ls)
        torch.cuda.synchronize()
        err = ((out - ref).abs().amax() / ref.abs().amax()).item()
        bad = int(((out - ref).abs() > 1e-3 * ref.abs().amax()).sum())
        print(f"shape {(B,C,T_y,T_x)} impl {impl}: rel err {err:.3e}  bad cells {bad}", flush=True)
        if impl == 1 and err > 1e-5:
            d = (out - ref).abs()
            idx = torch.nonzero(d > 1e-3 * ref.abs().amax())[:8].tolist()
            print("   first bad:", idx, out.flatten()[:4].tolist(), ref.flatten()[:4].tolist())
L.mas_set_neg_cent_impl(-1)
