"""CPU tests: the oracle (oracle/) against the reference's golden vectors and compiled code.

The oracle is the checker for the GPU parity tests, so it is pinned here first:
  * tests/golden/mas_golden.npz / synth_golden.npz were produced by the REAL reference
    (tests/golden/make_golden.py: its own wrapper + its own compiled Cython + its own model);
  * when oracle/_ref is present (built from /root/reference in place) the C port is also
    compared against the compiled reference on fresh random inputs.
"""
import numpy as np
import pytest
import torch

from helpers import check_path_properties, index_to_path, path_to_index, random_lengths, sha_path


def test_oracle_matches_reference_golden(oracle, mas_golden):
    assert len(mas_golden) >= 18
    for name, c in mas_golden.items():
        path = oracle.maximum_path_numpy(c["neg_cent"], c["t_ys"], c["t_xs"])
        assert sha_path(path) == c["sha256"], name
        np.testing.assert_array_equal(path_to_index(path), c["index"], err_msg=name)
        check_path_properties(path, c["t_ys"], c["t_xs"])


def test_appendix_a_known_answers(oracle, mas_golden):
    """SURVEY.md Appendix A (captured from the compiled reference during the survey)."""
    c = mas_golden["kat_zeros_6x3"]
    assert c["index"][0].tolist() == [0, 1, 2, 2, 2, 2]                 # ties stay, forced diagonal
    assert mas_golden["kat_square_5x5"]["index"][0].tolist() == [0, 1, 2, 3, 4]
    assert mas_golden["kat_tx1_4x1"]["index"][0].tolist() == [0, 0, 0, 0]
    k4 = mas_golden["kat_mod13_ragged"]["index"]
    assert k4[0].tolist() == [0, 0, 0, 1, 1, 2, 2, 2, 2, 3, 3, 4]
    assert k4[1].tolist() == [0, 0, 0, 0, 1, 1, 2, 2, 3, -1, -1, -1]
    assert k4[2].tolist() == [0, 0, 0, 1, 1, 2, 2, -1, -1, -1, -1, -1]
    k5 = mas_golden["kat_mod29_ragged"]
    assert k5["sha256"] == "fa4f108ba9b56b19f13f9213964b1a56a33065591db48c40e6f566b7924060ff"
    p5 = index_to_path(k5["index"].astype(np.int32), 16)
    assert p5[0].sum(0).tolist() == [1, 1, 1, 6, 1, 2, 8, 2, 1, 3, 1, 1, 12, 3, 7, 14]
    k6 = index_to_path(mas_golden["kat_hand_4x3"]["index"].astype(np.int32), 3)
    assert k6[0].tolist() == [[1, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]]


def test_pure_python_port_agrees(oracle, mas_golden):
    for name in ["kat_zeros_6x3", "kat_mod13_ragged", "kat_mod29_ragged", "kat_hand_4x3", "rand_3x40x31",
                 "ties_int_3x90x40", "huge_2x60x20"]:
        c = mas_golden[name]
        path = oracle.maximum_path_py(c["neg_cent"], c["t_ys"], c["t_xs"])
        assert sha_path(path) == c["sha256"], name


@pytest.mark.parametrize("kind", ["stock", "omp"])
def test_c_port_vs_compiled_reference(oracle, kind):
    core = oracle.load_ref_core(kind)
    if core is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    rng = np.random.default_rng(99)
    for (B, T_y, T_x) in [(4, 64, 17), (3, 257, 96), (2, 512, 192), (5, 33, 33), (2, 700, 1)]:
        nc = (rng.standard_normal((B, T_y, T_x)) * 2 - 3).astype(np.float32)
        t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
        ours = oracle.maximum_path_numpy(nc, t_ys, t_xs)
        ref = oracle.maximum_path_numpy(nc, t_ys, t_xs, core=core)
        np.testing.assert_array_equal(ours, ref)
        check_path_properties(ours, t_ys, t_xs)
    # integer-valued inputs (dense ties) and values below the -1e9 sentinel
    nc = rng.integers(-3, 4, size=(3, 120, 50)).astype(np.float32)
    t_ys, t_xs = random_lengths(rng, 3, 120, 50)
    np.testing.assert_array_equal(oracle.maximum_path_numpy(nc, t_ys, t_xs),
                                  oracle.maximum_path_numpy(nc, t_ys, t_xs, core=core))
    nc = (rng.standard_normal((2, 80, 30)) * 5e8 - 1e9).astype(np.float32)
    t_ys, t_xs = random_lengths(rng, 2, 80, 30)
    np.testing.assert_array_equal(oracle.maximum_path_numpy(nc, t_ys, t_xs),
                                  oracle.maximum_path_numpy(nc, t_ys, t_xs, core=core))


def test_wrapper_contract(oracle):
    """dtype/device follow neg_cent, input not mutated, bool mask accepted, non-contiguous rejected
    (reference behaviour recorded in SURVEY.md Appendix A)."""
    rng = np.random.default_rng(3)
    nc = torch.from_numpy(rng.standard_normal((2, 30, 10)).astype(np.float32))
    keep = nc.clone()
    t_ys, t_xs = torch.tensor([30, 21]), torch.tensor([10, 7])
    mask = oracle.attn_mask(t_xs, t_ys, 10, 30)
    out = oracle.maximum_path(nc, mask.float())
    assert out.dtype == torch.float32 and out.shape == nc.shape and torch.equal(nc, keep)
    out_b = oracle.maximum_path(nc, mask)            # bool mask
    assert torch.equal(out, out_b)
    out_h = oracle.maximum_path(nc.half(), mask.float())
    assert out_h.dtype == torch.float16
    out_d = oracle.maximum_path(nc.double(), mask.float())
    assert out_d.dtype == torch.float64 and torch.equal(out_d.float(), out)
    check_path_properties(out.numpy().astype(np.int8), t_ys.numpy(), t_xs.numpy())
    with pytest.raises(ValueError):
        oracle.maximum_path_c(np.zeros((2, 10, 30), np.int32).transpose(0, 2, 1),
                              np.zeros((2, 30, 10), np.float32), np.array([30, 21], np.int32),
                              np.array([10, 7], np.int32))


def test_neg_cent_expansion_vs_closed_form(oracle):
    g = torch.Generator().manual_seed(5)
    B, C, T_y, T_x = 2, 192, 70, 24
    z = torch.randn(B, C, T_y, generator=g)
    m = torch.randn(B, C, T_x, generator=g)
    ls = torch.randn(B, C, T_x, generator=g) * 0.3
    nc32 = oracle.neg_cent_torch(z, m, ls)
    nc64 = oracle.neg_cent_f64(z.numpy(), m.numpy(), ls.numpy())
    scale = np.abs(nc64).max()
    assert nc32.dtype == torch.float32 and tuple(nc32.shape) == (B, T_y, T_x)
    assert np.abs(nc32.numpy() - nc64).max() <= 1e-5 * scale
    nc64_t = oracle.neg_cent_torch(z.double(), m.double(), ls.double()).numpy()
    assert np.abs(nc64_t - nc64).max() <= 1e-11 * scale


def test_oracle_vs_real_model_capture(oracle, synth_golden):
    """One forward of the real SynthesizerTrn (config_cje.yaml, random init, CPU fp32)."""
    s = synth_golden
    z, m, ls = (torch.from_numpy(s[k]) for k in ("z_p", "m_p", "logs_p"))
    nc = oracle.neg_cent_torch(z, m, ls)
    ref = torch.from_numpy(s["neg_cent"])
    assert (nc - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    mask = torch.from_numpy(s["x_mask"]).unsqueeze(2) * torch.from_numpy(s["spec_mask"]).unsqueeze(-1)
    path = oracle.maximum_path(ref, mask.squeeze(1))
    np.testing.assert_array_equal(path_to_index(path.numpy()), s["index"])
    np.testing.assert_array_equal(mask.squeeze(1).sum(1)[:, 0].numpy().astype(np.int32), s["y_lengths"])
    np.testing.assert_array_equal(mask.squeeze(1).sum(2)[:, 0].numpy().astype(np.int32), s["x_lengths"])


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_neg_cent_autocast_restatement_vs_torch_cpu_autocast(oracle, dtype):
    """the numpy restatement of SynthesizerTrn.py:223-232 as it evaluates
    under autocast, against the literal expression inside torch's CPU autocast (fp32 inputs: only the einsums
    change precision).  torch accumulates the products in fp32 in its own order, the restatement rounds the
    exact sum once, so a few einsum outputs may fall on the other side of a rounding boundary."""
    import math
    g = torch.Generator().manual_seed(3)
    B, C, T_y, T_x = 2, 192, 70, 33
    z = torch.randn(B, C, T_y, generator=g)
    m = torch.randn(B, C, T_x, generator=g)
    ls = torch.randn(B, C, T_x, generator=g) * 0.3
    with torch.no_grad(), torch.autocast("cpu", dtype=dtype):
        iv = torch.exp(-2 * ls)
        e2 = torch.einsum("bdt, bds -> bts", -0.5 * (z ** 2), iv)
        e3 = torch.einsum("bdt, bds -> bts", z, m * iv)
        want = (torch.sum(-0.5 * math.log(2 * math.pi) - ls, [1], keepdim=True) + e2 + e3
                + torch.sum(-0.5 * (m ** 2) * iv, [1], keepdim=True))
    assert e2.dtype == dtype and e3.dtype == dtype and want.dtype == torch.float32
    got = oracle.neg_cent_autocast_np(z.numpy(), m.numpy(), ls.numpy(), str(dtype).split(".")[1])
    want = want.numpy()
    scale = np.abs(want).max()
    d = np.abs(got - want)
    eps = torch.finfo(dtype).eps
    assert d.max() <= eps * (e2.float().abs().max().item() + e3.float().abs().max().item()) + 1e-5 * scale
    assert (d <= 2e-6 * scale).mean() >= 0.99
    fp32 = oracle.neg_cent_torch(z, m, ls).numpy()              # and the mode is not the fp32 formulation
    assert (np.abs(fp32 - want) <= 2e-6 * scale).mean() < 0.5


def test_low_precision_rounding_helper_matches_torch(oracle):
    """oracle._round_to (nearest-even fp32 -> fp16 / bf16 -> fp32) against torch's casts, ties and subnormals included."""
    rng = np.random.default_rng(5)
    a = np.concatenate([rng.standard_normal(20000).astype(np.float32) * 300,
                        rng.standard_normal(2000).astype(np.float32) * 1e-6,
                        np.array([0.0, -0.0, 1.0, 1.00390625, 1.01171875, 1.0 + 2.0 ** -11, 1.0 + 3 * 2.0 ** -11,
                                  65504.0, -65504.0, 6e-8, 3e-8], dtype=np.float32)])
    t = torch.from_numpy(a)
    np.testing.assert_array_equal(oracle._round_to(a, "bfloat16"), t.to(torch.bfloat16).float().numpy())
    np.testing.assert_array_equal(oracle._round_to(a, "float16"), t.to(torch.float16).float().numpy())


def test_near_tie_report_accepts_perturbation_sized_gaps_only(oracle):
    """tests/helpers.near_tie_report (the end-to-end gate's confinement check): tiny input perturbations may only move
    the walk at near-ties; a large perturbation is reported as a real disagreement."""
    from helpers import near_tie_report, path_to_index
    rng = np.random.default_rng(3)
    B, T_y, T_x = 4, 300, 60
    nc = (rng.standard_normal((B, T_y, T_x)) * 20 - 400).astype(np.float32)
    t_ys = np.array([300, 280, 200, 150], np.int32)
    t_xs = np.array([60, 50, 44, 30], np.int32)
    ref = path_to_index(oracle.maximum_path_numpy(nc, t_ys, t_xs))
    small = (nc.astype(np.float64) * (1 + 2e-6 * rng.standard_normal(nc.shape))).astype(np.float32)
    got = path_to_index(oracle.maximum_path_numpy(small, t_ys, t_xs))
    near_tie_report(oracle, nc, small, ref, got, t_ys, t_xs)          # must not raise
    assert near_tie_report(oracle, nc, nc, ref, ref, t_ys, t_xs) == []
    # a path that differs although the inputs are IDENTICAL is never a near-tie (bound = rounding only)
    other = path_to_index(oracle.maximum_path_numpy(nc + rng.standard_normal(nc.shape).astype(np.float32) * 30, t_ys, t_xs))
    assert (other != ref).any()
    with pytest.raises(AssertionError):
        near_tie_report(oracle, nc, nc, ref, other, t_ys, t_xs)
