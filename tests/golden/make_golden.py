"""Generate tests/golden/*.npz by running the REAL reference (not our oracle).

Run in the authoring container only (needs /root/reference):

    make -C oracle ref && python tests/golden/make_golden.py

What runs:
* the reference's own ``maximum_path`` wrapper (monotonic_align/__init__.py, executed in place
  from /root/reference -- nothing is copied) on top of its own Cython core compiled by
  oracle/Makefile into oracle/_ref/stock/;
* the reference's ``SynthesizerTrn.forward`` (config_cje.yaml model section, random init, CPU,
  fp32, autocast off) with hooks that capture the hot path's inputs (z_p, m_p, logs_p, masks),
  its ``neg_cent`` and the ``attn`` it returns.

Outputs (committed, small):
* mas_golden.npz   -- known-answer + seeded random cases: inputs, lengths, per-frame index,
                      SHA-256 of the int8 path.
* synth_golden.npz -- one B=3 forward of the real model.
"""
import hashlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REFERENCE = os.environ.get("REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import mas_oracle  # noqa: E402  (only for load_ref_core: the compiled reference)


def import_reference_wrapper():
    """Import /root/reference/monotonic_align as the package ``monotonic_align`` with its nested
    ``monotonic_align.monotonic_align.core`` bound to the compiled oracle/_ref/stock module."""
    core_fn = mas_oracle.load_ref_core("stock")
    assert core_fn is not None, "run `make -C oracle ref` first"
    core_mod = sys.modules.get("core") or types.ModuleType("core")
    core_mod.maximum_path_c = core_fn
    inner = types.ModuleType("monotonic_align.monotonic_align")
    inner.__path__ = []
    inner.core = core_mod
    sys.modules["monotonic_align.monotonic_align"] = inner
    sys.modules["monotonic_align.monotonic_align.core"] = core_mod
    spec = importlib.util.spec_from_file_location(
        "monotonic_align", os.path.join(REFERENCE, "monotonic_align", "__init__.py"),
        submodule_search_locations=[os.path.join(REFERENCE, "monotonic_align")])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules["monotonic_align"] = pkg
    spec.loader.exec_module(pkg)
    return pkg.maximum_path


def mask_from_lengths(t_ys, t_xs, T_y, T_x):
    ym = torch.arange(T_y)[None, :] < torch.as_tensor(t_ys)[:, None]
    xm = torch.arange(T_x)[None, :] < torch.as_tensor(t_xs)[:, None]
    return (ym[:, :, None] & xm[:, None, :]).float()


def path_index(path):
    p = path.numpy().astype(np.int32)
    idx = p.argmax(-1).astype(np.int16)
    idx[p.sum(-1) == 0] = -1
    return idx


def sha_path(path):
    return hashlib.sha256(path.numpy().astype(np.int8).tobytes()).hexdigest()


def mas_cases():
    """(name, neg_cent float32 [B,T_y,T_x], t_ys, t_xs)."""
    cases = []
    # Appendix-A style known-answer inputs (exact in fp32)
    cases.append(("kat_zeros_6x3", np.zeros((1, 6, 3), np.float32), [6], [3]))
    rng = np.random.default_rng(7)
    cases.append(("kat_square_5x5", rng.standard_normal((1, 5, 5)).astype(np.float32), [5], [5]))
    cases.append(("kat_tx1_4x1", rng.standard_normal((1, 4, 1)).astype(np.float32), [4], [1]))
    b, y, x = np.meshgrid(np.arange(3), np.arange(12), np.arange(5), indexing="ij")
    cases.append(("kat_mod13_ragged", (((31 * y + 17 * x + 7 * b) % 13) - 6).astype(np.float32),
                  [12, 9, 7], [5, 4, 3]))
    b, y, x = np.meshgrid(np.arange(4), np.arange(64), np.arange(16), indexing="ij")
    cases.append(("kat_mod29_ragged",
                  (0.25 * (((7 * y * y + 13 * x * x + 5 * x * y + 11 * b) % 29) - 14)).astype(np.float32),
                  [64, 50, 33, 16], [16, 16, 9, 16]))
    cases.append(("kat_hand_4x3", np.array([[[1, 5, 2], [2, 1, 7], [3, 9, 1], [0, 2, 4]]], np.float32),
                  [4], [3]))
    # seeded random sweeps: warp/lane boundary widths, t_x==t_y, t_y==t_x+1, dense ties, huge magnitudes
    rng = np.random.default_rng(20241018)
    for (B, T_y, T_x) in [(3, 40, 31), (3, 70, 32), (3, 70, 33), (2, 200, 64), (2, 130, 65),
                          (2, 97, 96), (2, 97, 97), (2, 300, 191), (2, 260, 193), (1, 128, 32)]:
        nc = rng.standard_normal((B, T_y, T_x)).astype(np.float32) * 3 - 5
        t_xs = rng.integers(max(1, T_x // 2), T_x + 1, size=B)
        t_ys = np.array([rng.integers(max(tx, T_y // 2), T_y + 1) for tx in t_xs])
        t_xs[0], t_ys[0] = T_x, T_y
        cases.append((f"rand_{B}x{T_y}x{T_x}", nc, t_ys.tolist(), t_xs.tolist()))
    ties = rng.integers(-2, 3, size=(3, 90, 40)).astype(np.float32)
    cases.append(("ties_int_3x90x40", ties, [90, 61, 40], [40, 33, 40]))
    huge = (rng.standard_normal((2, 60, 20)) * 4e8 - 8e8).astype(np.float32)
    cases.append(("huge_2x60x20", huge, [60, 45], [20, 11]))
    return cases


def make_mas_golden(maximum_path):
    out = {}
    names = []
    for name, nc, t_ys, t_xs in mas_cases():
        B, T_y, T_x = nc.shape
        nc_t = torch.from_numpy(nc.copy())
        path = maximum_path(nc_t, mask_from_lengths(t_ys, t_xs, T_y, T_x))
        assert path.dtype == torch.float32 and torch.equal(nc_t, torch.from_numpy(nc))
        names.append(name)
        out[f"{name}/neg_cent"] = nc
        out[f"{name}/t_ys"] = np.asarray(t_ys, np.int32)
        out[f"{name}/t_xs"] = np.asarray(t_xs, np.int32)
        out[f"{name}/index"] = path_index(path)
        out[f"{name}/sha256"] = np.frombuffer(bytes.fromhex(sha_path(path)), dtype=np.uint8)
        print(f"{name:24s} sum={int(path.sum())} sha={sha_path(path)[:16]}")
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "mas_golden.npz"), **out)


def make_synth_golden(maximum_path):
    import yaml

    sys.path.insert(0, REFERENCE)
    import SynthesizerTrn as ST  # the reference module (binds monotonic_align.maximum_path at import)

    with open(os.path.join(REFERENCE, "configs", "config_cje.yaml")) as f:
        cfg = yaml.safe_load(f)
    torch.manual_seed(1234)
    data, model = cfg["data"], cfg["model"]
    net = ST.SynthesizerTrn(
        71, data["filter_length"] // 2 + 1, cfg["train"]["segment_size"] // data["hop_length"],
        n_speakers=len(data["speakers"]), midi_start=data["midi_start"], midi_end=data["midi_end"],
        octave_range=data["octave_range"], **model)
    net.eval()

    cap = {}
    net.text_encoder.register_forward_hook(lambda m, i, o: cap.update(m_p=o[1].detach(), logs_p=o[2].detach(),
                                                                        x_mask=o[3].detach()))
    net.flow.register_forward_hook(lambda m, i, o: cap.update(z_p=o.detach()))
    net.posterior_encoder.register_forward_hook(lambda m, i, o: cap.update(spec_mask=o[3].detach()))

    def capturing(neg_cent, mask):
        cap["neg_cent"] = neg_cent.detach().clone()
        cap["mask"] = mask.detach().clone()
        out = maximum_path(neg_cent, mask)
        cap["attn"] = out.clone()
        return out

    ST.maximum_path = capturing

    B, T_x, T_y = 3, 44, 150
    x_lengths = torch.tensor([44, 37, 23])
    y_lengths = torch.tensor([150, 121, 96])
    x = torch.randint(1, 71, (B, T_x))
    tone = torch.randint(0, 4, (B, T_x))
    for i in range(B):
        x[i, x_lengths[i]:] = 0
        tone[i, x_lengths[i]:] = 0
    spec = torch.rand(B, data["filter_length"] // 2 + 1, T_y)
    n_ying = net.pitch_encoder.pre.in_channels if hasattr(net.pitch_encoder, "pre") else None
    ying_ch = n_ying or 80
    ying = torch.rand(B, ying_ch, T_y)
    for i in range(B):
        spec[i, :, y_lengths[i]:] = 0
        ying[i, :, y_lengths[i]:] = 0
    sid = torch.tensor([0, 2, 4])
    with torch.no_grad():
        net(x, tone, x_lengths, spec, y_lengths, ying, y_lengths, sid)
    for k, v in cap.items():
        print(k, tuple(v.shape), v.dtype)
    np.savez_compressed(
        os.path.join(HERE, "synth_golden.npz"),
        z_p=cap["z_p"].numpy(), m_p=cap["m_p"].numpy(), logs_p=cap["logs_p"].numpy(),
        x_mask=cap["x_mask"].numpy(), spec_mask=cap["spec_mask"].numpy(),
        neg_cent=cap["neg_cent"].numpy(), index=path_index(cap["attn"]),
        x_lengths=x_lengths.numpy().astype(np.int32), y_lengths=y_lengths.numpy().astype(np.int32))


if __name__ == "__main__":
    mp = import_reference_wrapper()
    make_mas_golden(mp)
    make_synth_golden(mp)
