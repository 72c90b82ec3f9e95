"""GPU parity tests of neg_cent (SynthesizerTrn.py:223-232) and of the end-to-end path.

Tolerances (BASELINE.json north_star / SURVEY.md 8d):
  * neg_cent: |ours - fp32 torch expression| <= 1e-5 * max|neg_cent| per utterance;
  * end-to-end: >= 99.9 % of path cells equal the oracle's, every differing frame a near-tie.
"""
import numpy as np
import math

import pytest
import torch

from helpers import near_tie_report, path_to_index, random_lengths

pytestmark = pytest.mark.gpu
REL_TOL = 1e-5


def _inputs(B, C, T_y, T_x, seed, t_ys=None, t_xs=None):
    g = torch.Generator(device="cuda").manual_seed(seed)
    z = torch.randn(B, C, T_y, generator=g, device="cuda")
    m = torch.randn(B, C, T_x, generator=g, device="cuda")
    ls = torch.randn(B, C, T_x, generator=g, device="cuda") * 0.3
    if t_ys is not None:   # zero the padded tails like TextEncoder.py:99 / PosteriorEncoder.py:66
        for b in range(B):
            z[b, :, int(t_ys[b]):] = 0
            m[b, :, int(t_xs[b]):] = 0
            ls[b, :, int(t_xs[b]):] = 0
    return z, m, ls


def _check(ours, ref):
    scale = ref.abs().amax(dim=(1, 2), keepdim=True)
    err = ((ours - ref).abs() / scale).max().item()
    assert err <= REL_TOL, f"relative error {err:.3e} > {REL_TOL}"
    return err


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("shape", [(2, 192, 128, 32), (3, 192, 150, 44), (2, 192, 1024, 192), (1, 192, 777, 257),
                                   (2, 80, 130, 70), (1, 192, 1536, 256), (1, 192, 300, 512)])
def test_neg_cent_vs_fp32_expression(oracle, impl, shape):
    import vits_b200
    vits_b200._lib.lib().mas_set_neg_cent_impl(impl)
    try:
        B, C, T_y, T_x = shape
        z, m, ls = _inputs(B, C, T_y, T_x, seed=sum(shape))
        ours = vits_b200.neg_cent(z, m, ls)
        assert ours.dtype == torch.float32 and tuple(ours.shape) == (B, T_y, T_x)
        ref32 = oracle.neg_cent_torch(z, m, ls)
        _check(ours, ref32)
        if T_y * T_x <= 200 * 300:
            ref64 = torch.from_numpy(oracle.neg_cent_f64(z.cpu().numpy(), m.cpu().numpy(), ls.cpu().numpy())).cuda()
            _check(ours.double(), ref64)
    finally:
        vits_b200._lib.lib().mas_set_neg_cent_impl(-1)


@pytest.mark.parametrize("shape", [(64, 192, 1024, 192), (32, 192, 1536, 256), (8, 192, 4096, 512)], ids=["c2", "c3", "c4"])
def test_neg_cent_full_config_shapes(oracle, shape):
    """BASELINE.json configs[1..3] at full size (c4: T_x = 512 = two column tiles, T_y = 4096 = 32 frame tiles),
    variable lengths with zeroed padding, automatic implementation (tcgen05)."""
    import vits_b200
    B, C, T_y, T_x = shape
    rng = np.random.default_rng(5)
    t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
    z, m, ls = _inputs(B, C, T_y, T_x, seed=T_y + T_x, t_ys=t_ys, t_xs=t_xs)
    ours = vits_b200.neg_cent(z, m, ls)
    _check(ours, oracle.neg_cent_torch(z, m, ls))


def test_neg_cent_real_model_capture(oracle, synth_golden):
    import vits_b200
    s = synth_golden
    z, m, ls = (torch.from_numpy(s[k]).cuda() for k in ("z_p", "m_p", "logs_p"))
    ours = vits_b200.neg_cent(z, m, ls)
    _check(ours, torch.from_numpy(s["neg_cent"]).cuda())


def test_neg_cent_ignores_autocast(oracle):
    import vits_b200
    z, m, ls = _inputs(2, 192, 96, 40, seed=3)
    with torch.autocast("cuda", dtype=torch.float16):
        ours = vits_b200.neg_cent(z.half(), m, ls)
    _check(ours, oracle.neg_cent_torch(z.half().float(), m, ls))


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(3, 192, 260, 90), (2, 192, 129, 65), (8, 192, 512, 192)])
def test_neg_cent_autocast_parity_mode(oracle, dtype, shape):
    """the contraction with the numerics the reference has AS TRAINED.  The
    target is the literal expression (SynthesizerTrn.py:223-232) run under torch.autocast on the GPU: both
    einsums on low-precision operands with low-precision outputs, everything else fp32.  cuBLAS and our kernel
    add the fp32 partial products in a different order, so an einsum output may land on the other side of a
    rounding boundary: nearly all elements must be EQUAL, none further away than one ulp per einsum term, and
    the fp32 formulation (the default mode) must be clearly distinguishable from it."""
    import vits_b200
    z, m, ls = _inputs(*shape, seed=17)
    with torch.no_grad(), torch.autocast("cuda", dtype=dtype):   # oracle.neg_cent_torch switches autocast off: spell it out
        inv_var = torch.exp(-2 * ls)                                                        # :223
        want = (torch.sum(-0.5 * math.log(2 * math.pi) - ls, [1], keepdim=True)             # :225
                + torch.einsum("bdt, bds -> bts", -0.5 * (z ** 2), inv_var)                 # :227
                + torch.einsum("bdt, bds -> bts", z, m * inv_var)                           # :229
                + torch.sum(-0.5 * (m ** 2) * inv_var, [1], keepdim=True))                  # :231, :232
    assert want.dtype == torch.float32
    got = vits_b200.neg_cent(z, m, ls, autocast_dtype=dtype)
    assert got.dtype == torch.float32 and got.shape == want.shape
    eps = torch.finfo(dtype).eps
    with torch.autocast("cuda", dtype=dtype):   # magnitude of the two rounded terms
        s = torch.exp(-2 * ls)
        t2 = torch.einsum('bdt,bds->bts', -0.5 * z ** 2, s).float().abs().max().item()
        t3 = torch.einsum('bdt,bds->bts', z, m * s).float().abs().max().item()
    bound = eps * (t2 + t3) + 1e-4 * want.abs().max().item()
    diff = (got - want).abs()
    assert diff.max().item() <= bound, (diff.max().item(), bound)
    equal = (diff <= 2e-6 * want.abs().max()).float().mean().item()   # fp32 sums of terms 1 and 4 may differ in the last bit
    assert equal >= 0.9, equal
    fp32_mode = vits_b200.neg_cent(z, m, ls)
    assert ((fp32_mode - want).abs() <= 2e-6 * want.abs().max()).float().mean().item() < 0.5


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_neg_cent_autocast_low_precision_stats(oracle, dtype):
    """As in training: TextEncoder.proj hands m_p / logs_p over in the autocast dtype (TextEncoder.py:101-104), so
    the elementwise part of :225 is rounded to it.  Three-way: kernel vs the literal expression under CUDA
    autocast vs the numpy restatement (which pins the oracle's stats_lowp variant, see its docstring)."""
    import vits_b200
    z, m, ls = _inputs(3, 192, 200, 77, seed=23)
    m, ls = m.to(dtype), ls.to(dtype)
    with torch.no_grad(), torch.autocast("cuda", dtype=dtype):
        inv_var = torch.exp(-2 * ls)
        assert inv_var.dtype == torch.float32
        e1 = -0.5 * math.log(2 * math.pi) - ls
        assert e1.dtype == dtype
        want = (torch.sum(e1, [1], keepdim=True)
                + torch.einsum("bdt, bds -> bts", -0.5 * (z ** 2), inv_var)
                + torch.einsum("bdt, bds -> bts", z, m * inv_var)
                + torch.sum(-0.5 * (m ** 2) * inv_var, [1], keepdim=True))
    assert want.dtype == torch.float32
    got = vits_b200.neg_cent(z, m, ls, autocast_dtype=dtype)
    rest = torch.from_numpy(oracle.neg_cent_autocast_np(z.cpu().numpy(), m.float().cpu().numpy(), ls.float().cpu().numpy(),
                                                        str(dtype).split(".")[1], stats_lowp=True)).cuda()
    scale = want.abs().max()
    for a, b_ in ((got, want), (got, rest), (rest, want)):
        assert ((a - b_).abs() <= 2e-6 * scale).float().mean().item() >= 0.9
    # without the flag (stats widened by the caller beforehand) term 1 differs visibly
    plain = vits_b200.neg_cent(z, m.float(), ls.float(), autocast_dtype=dtype)
    assert ((plain - want).abs() <= 2e-6 * scale).float().mean().item() < 0.9


def test_neg_cent_autocast_mode_rejects_other_types():
    import vits_b200
    z, m, ls = _inputs(1, 192, 64, 32, seed=1)
    with pytest.raises(ValueError):
        vits_b200.neg_cent(z, m, ls, autocast_dtype=torch.float64)


@pytest.mark.parametrize("shape", [(8, 192, 400, 100), (64, 192, 1024, 192)])
def test_end_to_end_path_agreement(oracle, shape):
    """z_p, m_p, logs_p -> neg_cent (ours) -> path (ours)  vs  oracle neg_cent -> oracle path."""
    import vits_b200
    B, C, T_y, T_x = shape
    rng = np.random.default_rng(17)
    t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
    z, m, ls = _inputs(B, C, T_y, T_x, seed=29, t_ys=t_ys, t_xs=t_xs)
    nc_ours = vits_b200.neg_cent(z, m, ls)
    path_ours = vits_b200.maximum_path_from_lengths(nc_ours, torch.as_tensor(t_ys), torch.as_tensor(t_xs))
    nc_ref = oracle.neg_cent_torch(z, m, ls)
    want = oracle.maximum_path_numpy(nc_ref.cpu().numpy(), t_ys, t_xs)
    got = path_ours.cpu().numpy().astype(np.int32)
    valid = sum(int(a) * int(b) for a, b in zip(t_ys, t_xs))
    differing = int((got != want).sum())
    assert differing <= 1e-3 * valid, f"{differing} of {valid} cells differ"
    # ... and wherever the two walks part, the reference's own value table shows a near-tie (core.pyx:32)
    parts = near_tie_report(oracle, nc_ref.cpu().numpy(), nc_ours.cpu().numpy(), path_to_index(want), path_to_index(got),
                            t_ys, t_xs)
    assert (differing == 0) == (len(parts) == 0)
    # and on the *same* neg_cent the paths are bit-identical
    same = vits_b200.maximum_path_from_lengths(nc_ref, torch.as_tensor(t_ys), torch.as_tensor(t_xs))
    np.testing.assert_array_equal(same.cpu().numpy().astype(np.int32), want)


def test_maximum_path_from_stats(oracle):
    """Both stages in one call, lengths instead of a mask: equals maximum_path(neg_cent(...), attn_mask)."""
    import vits_b200
    g = torch.Generator(device="cuda").manual_seed(11)
    B, C, T_y, T_x = 3, 192, 260, 90
    z = torch.randn(B, C, T_y, generator=g, device="cuda")
    m = torch.randn(B, C, T_x, generator=g, device="cuda")
    ls = torch.randn(B, C, T_x, generator=g, device="cuda") * 0.3
    y_len = torch.tensor([260, 200, 131], device="cuda")
    x_len = torch.tensor([90, 71, 45], device="cuda")
    mask = oracle.attn_mask(x_len, y_len, T_x, T_y, torch.float32)
    want = vits_b200.maximum_path(vits_b200.neg_cent(z, m, ls), mask)
    got = vits_b200.maximum_path_from_stats(z, m, ls, x_len, y_len)
    assert torch.equal(got, want)
    idx = vits_b200.maximum_path_from_stats(z, m, ls, x_len, y_len, index=True)
    assert idx.dtype == torch.int32 and torch.equal(torch.nn.functional.one_hot(idx.clamp_min(0).long(), T_x).float() * (idx >= 0)[..., None], want)


@pytest.mark.parametrize("shape", [(2, 192, 130, 40), (3, 192, 260, 90), (4, 80, 300, 64), (5, 192, 700, 192),
                                   (3, 192, 520, 300), (2, 192, 1100, 256)])
@pytest.mark.parametrize("ragged", [True, False])
def test_streamed_stats_to_path_is_bit_identical(oracle, shape, ragged):
    """mas_stats_to_path (SynthesizerTrn.py:223-235 in one pass, the contraction streamed into the search through an
    L2-resident ring) must give exactly the path of mas_neg_cent followed by mas_maximum_path, and that path must be the
    oracle's on the identical neg_cent."""
    import vits_b200
    B, C, T_y, T_x = shape
    rng = np.random.default_rng(sum(shape))
    if ragged:
        t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
    else:
        t_ys, t_xs = np.full(B, T_y, np.int32), np.full(B, T_x, np.int32)
    z, m, ls = _inputs(B, C, T_y, T_x, seed=3, t_ys=t_ys, t_xs=t_xs)
    ty, tx = torch.as_tensor(t_ys).cuda(), torch.as_tensor(t_xs).cuda()
    got = vits_b200.maximum_path_from_stats(z, m, ls, tx, ty, streamed=True)
    idx = vits_b200.maximum_path_from_stats(z, m, ls, tx, ty, index=True, streamed=True)
    nc = vits_b200.neg_cent(z, m, ls)
    want = oracle.maximum_path_numpy(nc.cpu().numpy(), t_ys, t_xs)
    np.testing.assert_array_equal(got.cpu().numpy().astype(np.int32), want)
    np.testing.assert_array_equal(idx.cpu().numpy(), path_to_index(want))
    assert got.dtype == torch.float32 and vits_b200.status_nosync() == 0


def test_streamed_stats_to_path_bad_lengths_and_fallback(oracle):
    import vits_b200
    from vits_b200 import _lib
    B, C, T_y, T_x = 4, 192, 200, 64
    t_ys = np.array([200, 30, 150, 0], np.int32)
    t_xs = np.array([64, 50, 40, 10], np.int32)          # utterance 1: t_x > t_y, utterance 3: empty
    z, m, ls = _inputs(B, C, T_y, T_x, seed=5)
    vits_b200.status_nosync(reset=True)
    got = vits_b200.maximum_path_from_stats(z, m, ls, torch.as_tensor(t_xs), torch.as_tensor(t_ys), streamed=True)
    torch.cuda.synchronize()
    nc = vits_b200.neg_cent(z, m, ls).cpu().numpy()
    ok = [0, 2]
    want = oracle.maximum_path_numpy(nc[ok], t_ys[ok], t_xs[ok])
    np.testing.assert_array_equal(got[ok].cpu().numpy().astype(np.int32), want)
    assert int(got[[1, 3]].abs().sum().item()) == 0
    bits = vits_b200.status_nosync(reset=True)
    assert bits & _lib.MAS_STATUS_TX_GT_TY and bits & _lib.MAS_STATUS_EMPTY and not bits & _lib.MAS_STATUS_TIMEOUT
    # a shape the streamed form does not cover falls back to the two-call path (and says so when forced)
    z2, m2, ls2 = _inputs(1, 192, 700, 600, seed=6)
    ty2, tx2 = torch.tensor([700]), torch.tensor([600])
    with pytest.raises(_lib.MasError):
        vits_b200.maximum_path_from_stats(z2, m2, ls2, tx2, ty2, streamed=True)
    auto = vits_b200.maximum_path_from_stats(z2, m2, ls2, tx2, ty2)
    want2 = oracle.maximum_path_numpy(vits_b200.neg_cent(z2, m2, ls2).cpu().numpy(), ty2.numpy(), tx2.numpy())
    np.testing.assert_array_equal(auto.cpu().numpy().astype(np.int32), want2)
