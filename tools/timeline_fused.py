"""Phase timeline of one streamed stats->path call (globaltimer stamps written by the kernels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vits_b200
from vits_b200 import _lib
from bench import WORKLOADS, make_lengths
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
B, T_y, T_x = WORKLOADS[wl]
L = _lib.lib()
names = ["dp_first_start", "dp_last_frame_done", "dp_last_end", "bt_first_start", "bt_last_end", "gemm_first_start", "gemm_last_tile", "lengths_known"]
for ragged in (True, False):
    t_ys, t_xs = make_lengths(np.random.default_rng(1234), B, T_y, T_x, ragged)
    g = torch.Generator(device="cuda").manual_seed(1)
    sets = [(torch.randn(B, 192, T_y, generator=g, device="cuda"), torch.randn(B, 192, T_x, generator=g, device="cuda"),
             torch.randn(B, 192, T_x, generator=g, device="cuda") * 0.3) for _ in range(3)]
    ty, tx = torch.as_tensor(t_ys).cuda(), torch.as_tensor(t_xs).cuda()
    tl = torch.zeros(16, dtype=torch.int64, device="cuda")
    runs = []
    for i in range(6):
        tl.zero_(); tl[0] = tl[3] = tl[5] = -1
        torch.cuda.synchronize()
        L.mas_set_timeline(tl.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        vits_b200.maximum_path_from_stats(*sets[i % 3], tx, ty, streamed=True)
        e1.record()
        torch.cuda.synchronize()
        L.mas_set_timeline(None)
        v = tl.cpu().numpy().astype(np.uint64)
        t0 = min(int(v[5]), int(v[0]))
        runs.append([(int(x) - t0) / 1e3 if int(x) not in (0, 2 ** 64 - 1) else float("nan") for x in v] + [e0.elapsed_time(e1) * 1e3])
    med = np.nanmedian(np.array(runs[1:]), axis=0)
    print(wl, "ragged" if ragged else "full", {n: round(float(m), 1) for n, m in zip(names + ["event_us(eager)"], med)})
