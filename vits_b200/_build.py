"""Build libvits_mas.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).
Each .cu is compiled to an object in parallel, then linked."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OBJ = os.path.join(_HERE, "build")
LIB_PATH = os.path.join(_HERE, "libvits_mas.so")
SOURCES = ["mas_path.cu", "mas_fwd_k1.cu", "mas_fwd_k2.cu", "mas_fwd_k3.cu", "mas_fwd_k4.cu", "mas_fwd_k6.cu", "mas_fwd_k8.cu", "mas_dp_k1.cu", "mas_dp_k2.cu", "mas_dp_k4.cu", "mas_dp2_k2.cu", "mas_neg_cent.cu",
           "mas_neg_cent_tc.cu", "mas_fused.cu", "mas_consumers.cu", "mas_api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math",
    # Nothing but the extern "C" surface is exported, and function-local statics of inline/template functions are
    # ordinary weak symbols instead of STB_GNU_UNIQUE: a UNIQUE symbol is shared by every copy of the library in a
    # process, and a second copy (e.g. the Cython binding's dependency next to a ctypes load of a rebuilt file) then saw
    # "shared-memory opt-in already applied" flags that belonged to the OTHER copy's kernels -> launches failed with
    # cudaErrorInvalidValue (r02 finding).
    "-Xcompiler", "-fvisibility=hidden", "-Xcompiler", "-fno-gnu-unique",
    "--fmad=true",   # FMA contraction is fine for neg_cent; the DP has no multiply to contract
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libvits_mas.so cannot be built (no CPU fallback exists)")


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(os.path.dirname(_HERE), "include", "vits_mas.h")]


STAMP = LIB_PATH + ".srchash"


def source_hash() -> str:
    """Digest of every source the library is built from (content, not mtimes: a snapshot copied to another
    box keeps its contents but not necessarily its timestamps)."""
    h = hashlib.sha256()
    for d in sorted(_deps()):
        if d.endswith((".cu", ".cuh", ".h")):
            h.update(os.path.basename(d).encode())
            with open(d, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build() -> bool:
    """True when libvits_mas.so is missing or was built from other sources than the ones in csrc/ now."""
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_hash()


def _compile_one(nvcc, src, obj, verbose):
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    out = subprocess.run(cmd, capture_output=True, text=True)
    return cmd, out


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    headers_mtime = max(os.path.getmtime(d) for d in _deps() if not d.endswith(".cu"))
    jobs = []
    for s in srcs:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), headers_mtime)
        stale = stale or not os.path.exists(STAMP)
        jobs.append((src, obj, stale))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(lambda j: _compile_one(nvcc, j[0], j[1], verbose) if j[2] else (None, None), jobs))
    for cmd, out in results:
        if out is None:
            continue
        if verbose or out.returncode != 0:
            print(" ".join(cmd))
            print(out.stdout)
            print(out.stderr)
        if out.returncode != 0:
            raise RuntimeError("nvcc failed building libvits_mas.so")
    link = [nvcc, "-shared", "-o", LIB_PATH] + [j[1] for j in jobs]
    out = subprocess.run(link, capture_output=True, text=True)
    if out.returncode != 0:
        print(out.stdout, out.stderr)
        raise RuntimeError("linking libvits_mas.so failed")
    with open(STAMP, "w") as f:
        f.write(source_hash())
    return LIB_PATH


if __name__ == "__main__":
    build(force=True, verbose=True)
