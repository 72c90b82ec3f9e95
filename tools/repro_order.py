"""Order dependence seen in r02: host-buffer entry first, then the device entry -> cudaErrorInvalidValue."""
import subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
import vits_b200
from vits_b200 import _lib
L = _lib.lib()
shapes = %s
for (B, T_y, T_x) in shapes:
    nc = torch.randn(B, T_y, T_x)
    out = vits_b200.maximum_path_from_lengths(nc, torch.full((B,), T_y), torch.full((B,), min(T_x, T_y)))
    print("host entry ok", (B, T_y, T_x), int(out.sum()))
try:
    nc = torch.randn(6, 330, 64, device="cuda")
    out = vits_b200.maximum_path_from_lengths(nc, torch.full((6,), 330), torch.full((6,), 64))
    torch.cuda.synchronize()
    print("device entry ok", int(out.sum()))
except Exception as e:
    print("device entry FAILED:", e)
'''
for shapes in ([(1, 6, 3), (1, 5, 5), (1, 4, 1), (2, 200, 64)], [(2, 200, 64)], [(1, 6, 3)], [(1, 4, 1)], [(3, 70, 32)], []):
    print("=== host-entry shapes first:", shapes, flush=True)
    r = subprocess.run([sys.executable, "-c", CODE % (ROOT, repr(shapes))], capture_output=True, text=True)
    print(r.stdout, r.stderr[-600:] if r.returncode else "", flush=True)
