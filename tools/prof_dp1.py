import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vits_b200
from vits_b200 import _lib
B, T_y, T_x = 64, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 64
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
K = int(sys.argv[3]) if len(sys.argv) > 3 else 0
L = _lib.lib()
ty = torch.full((B,), T_y, dtype=torch.int32).cuda(); tx = torch.full((B,), T_x, dtype=torch.int32).cuda()
g = torch.Generator(device='cuda').manual_seed(1)
bufs = [torch.randn(B, T_y, T_x, generator=g, device='cuda') * 20 - 400 for _ in range(3)]
L.mas_set_debug_kernels(1); L.mas_set_tuning(K, mode, 0, 0)
tl = torch.zeros(16, dtype=torch.int64, device='cuda')
res = []
for i in range(4):
    tl.zero_(); tl[0] = -1; torch.cuda.synchronize(); L.mas_set_timeline(tl.data_ptr())
    out = vits_b200.maximum_path_from_lengths(bufs[i % 3], ty, tx)
    torch.cuda.synchronize(); L.mas_set_timeline(None)
    v = tl.cpu().numpy().astype(np.uint64); res.append((int(v[1]) - int(v[0])) / 1e3)
print(f"T_x={T_x} mode={mode} K={K} nohandoff={os.environ.get('MAS_DBG_NOHANDOFF')}: dp span us", res)
