"""neg_cent timing on one workload: reference torch expression vs our kernels (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import vits_b200
from oracle import mas_oracle
from bench import WORKLOADS
L = vits_b200._lib.lib()
wl = sys.argv[1] if len(sys.argv) > 1 else 'c2'
B, T_y, T_x = WORKLOADS[wl]; C = 192
g = torch.Generator(device='cuda').manual_seed(1)
sets = [(torch.randn(B, C, T_y, generator=g, device='cuda'), torch.randn(B, C, T_x, generator=g, device='cuda'),
         torch.randn(B, C, T_x, generator=g, device='cuda') * 0.3) for _ in range(4)]
def timeit(fn, graph=True, reps=10):
    for s in sets[:2]: fn(*s)
    torch.cuda.synchronize()
    if graph:
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            outs = [fn(*s) for s in sets]
        run = gr.replay
    else:
        def run():
            for s in sets: fn(*s)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * len(sets)) * 1e3
flops = 4.0 * B * T_y * T_x * C
bytes_alg = 4.0 * (B * C * T_y + 2 * B * C * T_x + B * T_y * T_x)
print(f"workload {wl}: {flops/1e9:.2f} algorithmic GFLOP, {bytes_alg/1e6:.1f} MB algorithmic")
t = timeit(mas_oracle.neg_cent_torch); print(f"torch fp32 expression (reference formulation, cuBLAS SGEMM + ATen): {t:8.1f} us")
torch.backends.cuda.matmul.allow_tf32 = True
t = timeit(mas_oracle.neg_cent_torch); print(f"torch expression with allow_tf32 (not parity-grade):               {t:8.1f} us")
torch.backends.cuda.matmul.allow_tf32 = False
for impl, name in ((0, "ours fp32 CUDA-core kernel"), (1, "ours tcgen05 split-bf16 (prep + GEMM)")):
    L.mas_set_neg_cent_impl(impl)
    t = timeit(vits_b200.neg_cent)
    print(f"{name:66s}: {t:8.1f} us  -> {flops/t/1e6:7.1f} algorithmic TFLOP/s, {bytes_alg/t/1e3:7.1f} GB/s")
L.mas_set_neg_cent_impl(-1)
y_len = torch.full((B,), T_y, device='cuda'); x_len = torch.full((B,), T_x, device='cuda')
def e2e(z, m, ls):
    return vits_b200.maximum_path_from_lengths(vits_b200.neg_cent(z, m, ls), y_len, x_len)
t = timeit(e2e); print(f"end to end z_p,m_p,logs_p -> path (neg_cent + maximum_path): {t:8.1f} us -> {B/t*1e6:,.0f} alignments/s")
