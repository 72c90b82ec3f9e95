"""Build an alternative libvits_mas.so with extra nvcc defines (same ABI), for experiments:
    python tools/build_variant.py trace -DMAS_TRACE          -> vits_b200/build_trace/libvits_mas_trace.so
Use it with VITS_MAS_LIB=<path> (vits_b200/_lib.py)."""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vits_b200 import _build

def main():
    name, defs = sys.argv[1], sys.argv[2:]
    csrc = _build.CSRC
    if "--csrc" in defs:   # build from another copy of csrc/ (e.g. an older commit checked out under /tmp)
        i = defs.index("--csrc"); csrc = defs[i + 1]; defs = defs[:i] + defs[i + 2:]
    obj = os.path.join(ROOT, "vits_b200", "build_" + name)
    os.makedirs(obj, exist_ok=True)
    out = os.path.join(obj, f"libvits_mas_{name}.so")
    nvcc = _build._nvcc()
    def one(s):
        o = os.path.join(obj, s[:-3] + ".o")
        r = subprocess.run([nvcc] + _build.NVCC_FLAGS + defs + ["-c", "-o", o, os.path.join(csrc, s)], capture_output=True, text=True)
        if r.returncode: print(r.stdout, r.stderr); raise SystemExit(1)
        return o
    with ThreadPoolExecutor(8) as ex: objs = list(ex.map(one, _build.SOURCES))
    r = subprocess.run([nvcc, "-shared", "-o", out] + objs, capture_output=True, text=True)
    if r.returncode: print(r.stdout, r.stderr); raise SystemExit(1)
    for o in objs: os.remove(o)   # only the library travels to the GPU box
    print(out)
main()
