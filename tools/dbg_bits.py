"""Compare the forward kernel's decision bits with bits derived from a float32 numpy DP (debug)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vits_b200.monotonic_align as mp
L = mp._lib.lib()
B, T_y, T_x, mode = (int(a) for a in sys.argv[1:5])
SS = int(sys.argv[5]) if len(sys.argv) > 5 else 0
rng = np.random.default_rng(5)
nc = (rng.standard_normal((B, T_y, T_x)) * 3 - 4).astype(np.float32)
t_ys = np.full(B, T_y, np.int32); t_xs = np.full(B, T_x, np.int32)
NEG = np.float32(-1e9)
def ref_bits(v, ty, tx):
    """bit[y][x] = backtrack steps left when leaving frame y at column x (core.pyx:32), using the DP values."""
    val = v.copy(); bits = np.zeros((ty, tx), np.uint8)
    prev = np.full(tx, NEG, np.float32)
    for y in range(ty):
        stay = prev.copy(); 
        if y < tx: stay[y] = NEG
        step = np.empty(tx, np.float32); step[1:] = prev[:-1]; step[0] = np.float32(0.) if y == 0 else NEG
        b = (stay < step).astype(np.uint8)
        if y < tx: b[y] = 1
        b[0] = 0
        bits[y] = b
        prev = (val[y] + np.maximum(step, stay)).astype(np.float32)
    return bits
L.mas_set_tuning(0, mode, SS, 1); L.mas_set_tuning2(0, 0); L.mas_set_debug_kernels(1)
ncd = torch.from_numpy(nc).cuda()
nbytes = L.mas_maximum_path_scratch_bytes(B, T_y, T_x)
scratch = torch.zeros(nbytes, dtype=torch.uint8, device='cuda')
idx = torch.zeros(B, T_y, dtype=torch.int32, device='cuda')
ty = torch.as_tensor(t_ys).cuda(); tx = torch.as_tensor(t_xs).cuda()
rc = L.mas_maximum_path(ncd.data_ptr(), ty.data_ptr(), tx.data_ptr(), None, 0, 0, 0, 0, None, 0, idx.data_ptr(), scratch.data_ptr(), nbytes, B, T_y, T_x, None)
torch.cuda.synchronize(); assert rc == 0, rc
G = (T_y + 31) // 32
K = 2; W = (T_x + 63) // 64; TXP = W * 64
up = lambda v: (v + 255) & ~255
off = up(256 + B * 8); off = up(off + B * 8); off = up(off + B * T_y * 4)
words = scratch[off: off + B * G * TXP * 4].view(torch.int32).cpu().numpy().astype(np.uint32).reshape(B, G, TXP)
for b in range(B):
    rb = ref_bits(nc[b], T_y, T_x)
    got = np.zeros((T_y, T_x), np.uint8)
    for y in range(T_y):
        got[y] = (words[b, y >> 5, :T_x] >> (31 - (y & 31))) & 1
    # only in-band cells matter: x <= y and x >= t_x + y - t_y
    ys, xs = np.nonzero(got != rb)
    keep = (xs <= ys) & (xs >= T_x + ys - T_y)
    ys, xs = ys[keep], xs[keep]
    print(f"utt {b}: {len(ys)} in-band bit mismatches", "" if not len(ys) else f"first (y={ys[0]}, x={xs[0]}) warp {xs[0]//64} lane {(xs[0]%64)//2}; ys {np.unique(ys)[:12]} xs {np.unique(xs)[:12]}")
L.mas_set_debug_kernels(7)
