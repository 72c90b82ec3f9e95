"""BASELINE.json configs[4]: the reference's own `SynthesizerTrn.forward` (unmodified, from baseline/_ref/vits) with the
alignment drop-in bound to the name it imports at SynthesizerTrn.py:16 and calls at :235.

  * same seeded model and batch through `forward` twice -- stock Cython `maximum_path` vs `vits_b200.maximum_path` --
    `attn` bit-identical, in fp32 and under autocast(fp16) (train_and_evaluate.py:55);
  * with the contraction replaced too (`maximum_path_from_stats`): >= 99.9 % of cells equal and every parting of the
    two walks a near-tie in the reference's value table (SURVEY.md 8d gate 3);
  * one full training iteration (bench_c5.train_step: D step + G step, GradScaler) with the drop-in leaves no status bit.
"""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import near_tie_report, path_to_index

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c5():
    import bench_c5
    if not os.path.isdir(bench_c5.REF_VITS):
        pytest.skip("baseline/_ref/vits missing (python tools/install_reference.py)")
    mods = bench_c5.load_reference()
    return bench_c5, mods, bench_c5.load_config()


def _forward(bench_c5, mods, cfg, arm, fp16, B=6, T_x=64, T_y=330):
    ST, AV, losses, commons, ref_mas = mods
    dev = torch.device("cuda", 0)
    net_g, _ = bench_c5.build_models(ST, AV, cfg, dev, seed=99)
    net_g.train()
    taps = bench_c5.Taps(ST, net_g, arm, ref_mas.maximum_path)
    taps.keep_inputs = True
    batch = bench_c5.synthetic_batch(cfg, B, T_x, T_y, dev, seed=7)
    taps.cap["y_lengths"] = batch["spec_lengths"]
    torch.manual_seed(4242)       # dropout / slicing draw the same numbers in every arm
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16, enabled=fp16):
        out = net_g(batch["x"], batch["tone"], batch["x_lengths"], batch["spec"], batch["spec_lengths"], batch["ying"],
                    batch["spec_lengths"], batch["sid"])
    torch.cuda.synchronize()
    return out[2], taps.keep, batch


@pytest.mark.parametrize("fp16", [False, True], ids=["fp32", "autocast_fp16"])
def test_forward_attn_bit_identical_with_dropin(c5, fp16):
    bench_c5, mods, cfg = c5
    attn_s, keep_s, batch = _forward(bench_c5, mods, cfg, "stock", fp16)
    attn_d, keep_d, _ = _forward(bench_c5, mods, cfg, "dropin", fp16)
    assert keep_s["neg_cent"].dtype == torch.float32            # promoted by the fp32 terms (:225,:231) even under autocast
    assert torch.equal(keep_s["neg_cent"], keep_d["neg_cent"]), "both arms must search the identical neg_cent"
    assert attn_d.dtype == attn_s.dtype and attn_d.shape == attn_s.shape and attn_d.device == attn_s.device
    assert torch.equal(attn_s, attn_d), "attn differs between the stock Cython path and the drop-in"
    assert int(attn_d.sum().item()) == int(batch["spec_lengths"].sum().item())
    import vits_b200
    assert vits_b200.last_status() == 0 and vits_b200.status_nosync() == 0


def test_forward_with_contraction_replaced_too(c5, oracle):
    bench_c5, mods, cfg = c5
    attn_s, keep_s, batch = _forward(bench_c5, mods, cfg, "stock", False)
    attn_f, keep_f, _ = _forward(bench_c5, mods, cfg, "dropin_stats", False)
    import vits_b200
    t_ys, t_xs = batch["t_ys"], batch["t_xs"]
    want = attn_s[:, 0].cpu().numpy().astype(np.int32)
    got = attn_f[:, 0].cpu().numpy().astype(np.int32)
    valid = int(np.sum(t_ys.astype(np.int64) * t_xs))
    differing = int((want != got).sum())
    assert differing <= 1e-3 * valid, f"{differing} of {valid} cells differ"
    nc_ref = keep_s["neg_cent"].cpu().numpy()
    nc_got = vits_b200.neg_cent(keep_s["z_p"], keep_s["m_p"], keep_s["logs_p"]).cpu().numpy()
    scale = np.abs(nc_ref).max(axis=(1, 2), keepdims=True)
    assert (np.abs(nc_got - nc_ref) / scale).max() <= 1e-5
    parts = near_tie_report(oracle, nc_ref, nc_got, path_to_index(want), path_to_index(got), t_ys, t_xs)
    assert (differing == 0) == (len(parts) == 0)


def test_training_iteration_with_dropin_raises_no_status(c5):
    bench_c5, mods, cfg = c5
    import vits_b200
    dev = torch.device("cuda", 0)
    res = bench_c5.run_arm("dropin", mods, cfg, dev, 0, 1, 4, 48, 260, steps=2, warmup=1, fp16=True)
    assert res["status_bits_seen"] == 0
    assert 0.0 < res["mas_ms"] <= res["alignment_ms"] < res["step_ms"]
