"""Does the mask walk's memory traffic alone slow the call down?  Lengths GIVEN, and (mode 37) the idle warps walk the
mask's column 0 anyway and throw the sums away.  Raw C entry with both the lengths and the mask."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vits_b200
from vits_b200 import _lib
from bench import WORKLOADS, make_lengths
L = _lib.lib()
B, T_y, T_x = WORKLOADS['c2']
g = torch.Generator(device='cuda').manual_seed(1)
bufs = [torch.randn(B, T_y, T_x, generator=g, device='cuda') * 20 - 400 for _ in range(4)]
outs = [torch.empty(B, T_y, T_x, device='cuda') for _ in range(4)]
t_ys, t_xs = make_lengths(np.random.default_rng(1234), B, T_y, T_x, True)
ty, tx = torch.as_tensor(t_ys).cuda(), torch.as_tensor(t_xs).cuda()
mask = ((torch.arange(T_y, device='cuda')[None, :] < ty[:, None])[:, :, None] & (torch.arange(T_x, device='cuda')[None, :] < tx[:, None])[:, None, :]).float()
nbytes = int(L.mas_maximum_path_scratch_bytes(B, T_y, T_x))
scratch = torch.zeros(nbytes, dtype=torch.uint8, device='cuda')
stream = torch.cuda.current_stream().cuda_stream
def call(i, lens, msk):
    rc = L.mas_maximum_path(bufs[i].data_ptr(), ty.data_ptr() if lens else None, tx.data_ptr() if lens else None,
                            mask.data_ptr() if msk else None, _lib.MAS_F32, mask.stride(0), mask.stride(1), mask.stride(2),
                            outs[i].data_ptr(), _lib.MAS_F32, None, scratch.data_ptr(), scratch.numel(), B, T_y, T_x, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
for wf, lens, msk, what in ((33, True, False, "lengths given"), (33, False, True, "lengths from the mask"), (37, True, True, "lengths given + dummy walk of the mask"),
                            (33, True, False, "lengths given"), (37, True, True, "lengths given + dummy walk of the mask")):
    L.mas_set_tuning3(wf, 0, 0, 0)
    for i in range(2): call(i, lens, msk)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(8): call(i % 4, lens, msk)
    for _ in range(3): gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(25): gr.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{what:45s} {e0.elapsed_time(e1) * 1e3 / 200:7.2f} us/call", flush=True)
L.mas_set_tuning3(-1, 0, 0, 0)
