import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def mas_golden():
    z = np.load(os.path.join(GOLDEN, "mas_golden.npz"))
    cases = {}
    for name in z["names"]:
        name = str(name)
        cases[name] = dict(neg_cent=z[f"{name}/neg_cent"], t_ys=z[f"{name}/t_ys"], t_xs=z[f"{name}/t_xs"],
                           index=z[f"{name}/index"], sha256=bytes(z[f"{name}/sha256"]).hex())
    return cases


@pytest.fixture(scope="session")
def synth_golden():
    z = np.load(os.path.join(GOLDEN, "synth_golden.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def oracle():
    from oracle import mas_oracle
    mas_oracle.build()
    return mas_oracle
