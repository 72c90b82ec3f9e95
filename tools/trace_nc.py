import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vits_b200
from bench import WORKLOADS
B, T_y, T_x = WORKLOADS['c2']; C = 192
g = torch.Generator(device='cuda').manual_seed(1)
z = torch.randn(B, C, T_y, generator=g, device='cuda'); m = torch.randn(B, C, T_x, generator=g, device='cuda'); ls = torch.randn(B, C, T_x, generator=g, device='cuda') * 0.3
for i in range(2): vits_b200.neg_cent(z, m, ls)
torch.cuda.synchronize()
os.environ['MAS_NC_TRACE_DUMP'] = '1'
vits_b200.neg_cent(z, m, ls); torch.cuda.synchronize()
