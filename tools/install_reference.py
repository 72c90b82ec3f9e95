#!/usr/bin/env python
"""Install the UNMODIFIED reference modules that config 5 and the native-binding test need into the
git-ignored ``baseline/_ref/`` (it travels to the GPU box with the snapshot; ``/root/reference`` does not).

    python tools/install_reference.py            # needs /root/reference; a no-op message otherwise

What lands where (nothing here is product code, nothing is tracked by git):

  baseline/_ref/vits/                      the generator / discriminator / loss modules `SynthesizerTrn.forward`
                                           needs (they import with torch + numpy + scipy only),
                                           `configs/config_cje.yaml`, and the reference's `monotonic_align/`
                                           package with ITS OWN Cython core (the build oracle/Makefile makes
                                           from core.pyx, flags of its setup.py) in the nested directory its
                                           `__init__.py:4` imports from -- the stock arm of config 5.
  baseline/_ref/binding/monotonic_align/   the reference's `monotonic_align/__init__.py`, byte for byte, on top
                                           of OUR Cython binding (vits_b200/binding/core.pyx: same signature as
                                           core.pyx:38, body = one call into libvits_mas.so) -- the literal
                                           "reference-side binding" of INTEGRATION.md section 2.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys
import sysconfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("REFERENCE", "/root/reference")
DEST = os.path.join(ROOT, "baseline", "_ref")

MODULES = ["SynthesizerTrn.py", "Avocodo.py", "losses.py", "commons.py", "HiFiGANGenerator.py", "Pitch.py",
           "PosteriorEncoder.py", "ResidualCouplingBlock.py", "StochasticDurationPredictor.py", "TextEncoder.py",
           "YingDecoder.py", "WaveNet.py", "LayerNorm.py"]


def install_vits() -> str:
    dst = os.path.join(DEST, "vits")
    os.makedirs(os.path.join(dst, "configs"), exist_ok=True)
    for m in MODULES:
        shutil.copyfile(os.path.join(REFERENCE, m), os.path.join(dst, m))
    shutil.copyfile(os.path.join(REFERENCE, "configs", "config_cje.yaml"), os.path.join(dst, "configs", "config_cje.yaml"))
    pkg = os.path.join(dst, "monotonic_align")
    inner = os.path.join(pkg, "monotonic_align")
    os.makedirs(inner, exist_ok=True)
    shutil.copyfile(os.path.join(REFERENCE, "monotonic_align", "__init__.py"), os.path.join(pkg, "__init__.py"))
    # the reference's own compiled core (oracle/Makefile: cython on core.pyx in place, its setup.py's flags)
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True, capture_output=True)
    cores = glob.glob(os.path.join(ROOT, "oracle", "_ref", "stock", "core*.so"))
    assert cores, "oracle/_ref/stock/core*.so missing (make -C oracle ref)"
    shutil.copyfile(cores[0], os.path.join(inner, os.path.basename(cores[0])))
    return dst


def build_binding() -> str:
    """Cython-compile vits_b200/binding/core.pyx against libvits_mas.so and put it under the reference's
    unmodified monotonic_align/__init__.py."""
    import numpy
    pkg = os.path.join(DEST, "binding", "monotonic_align")
    inner = os.path.join(pkg, "monotonic_align")
    os.makedirs(inner, exist_ok=True)
    shutil.copyfile(os.path.join(REFERENCE, "monotonic_align", "__init__.py"), os.path.join(pkg, "__init__.py"))
    pyx = os.path.join(ROOT, "vits_b200", "binding", "core.pyx")
    c_out = os.path.join(DEST, "binding", "core.c")
    subprocess.run([sys.executable, "-m", "cython", "-3", "-I", os.path.join(ROOT, "include"), "-o", c_out, pyx], check=True)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    so = os.path.join(inner, "core" + ext)
    libdir = os.path.join(ROOT, "vits_b200")
    cflags = (sysconfig.get_config_var("CFLAGS") or "").split() + (sysconfig.get_config_var("CCSHARED") or "-fPIC").split()
    cmd = ["/usr/bin/gcc"] + cflags + ["-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(),
                                       "-I" + os.path.join(ROOT, "include"), "-shared", "-o", so, c_out,
                                       "-L" + libdir, "-lvits_mas", "-Wl,-rpath,$ORIGIN/../../../../../vits_b200"]
    subprocess.run(cmd, check=True)
    return so


def main() -> int:
    if not os.path.isdir(REFERENCE):
        print(f"install_reference: {REFERENCE} not present; keeping whatever baseline/_ref holds")
        return 0
    print("vits modules ->", install_vits())
    if os.path.exists(os.path.join(ROOT, "vits_b200", "libvits_mas.so")):
        print("native binding ->", build_binding())
    else:
        print("native binding skipped: build libvits_mas.so first (python -c 'import __graft_entry__ as g; g.build()')")
    return 0


if __name__ == "__main__":
    sys.exit(main())
