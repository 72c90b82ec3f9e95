"""Callers either side of the path: the oracle restatements against the reference's golden
vectors (CPU), and the CUDA kernels against the oracle (GPU, through the C ABI via the Python host side)."""
import os

import numpy as np
import pytest
import torch

from helpers import path_to_index, random_lengths

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "consumers_golden.npz")


def _golden():
    z = np.load(GOLDEN)
    return {str(n): (z[f"{n}/duration"], z[f"{n}/mask"], z[f"{n}/path"]) for n in z["names"]}


def test_oracle_generate_path_matches_reference_golden(oracle):
    for name, (dur, mask, path) in _golden().items():
        got = oracle.generate_path_torch(torch.from_numpy(dur), torch.from_numpy(mask)).numpy()
        np.testing.assert_array_equal(got, path, err_msg=name)
        # the path built from durations hands every token exactly its duration (inside the mask)
        np.testing.assert_array_equal(got.sum(2), dur * (mask.sum(2) > 0))


def test_oracle_consumers_are_consistent(oracle):
    rng = np.random.default_rng(3)
    B, T_y, T_x, C = 2, 30, 9, 5
    nc = rng.standard_normal((B, T_y, T_x)).astype(np.float32)
    t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
    path = oracle.maximum_path_numpy(nc, t_ys, t_xs).astype(np.float32)
    attn = torch.from_numpy(path).unsqueeze(1)
    w = oracle.durations_torch(attn)
    assert w.shape == (B, 1, T_x) and w.sum().item() == float(np.sum(t_ys))
    stat = torch.from_numpy(rng.standard_normal((B, C, T_x)).astype(np.float32))
    exp = oracle.expand_prior_torch(attn, stat).numpy()
    idx = path_to_index(path)
    for b in range(B):
        for y in range(T_y):
            want = stat[b, :, idx[b, y]].numpy() if idx[b, y] >= 0 else np.zeros(C, np.float32)
            np.testing.assert_array_equal(exp[b, :, y], want)
    # generate_path(durations of a path) reproduces the path: MAS output and duration-built paths agree
    mask = torch.from_numpy((path.sum(2, keepdims=True) > 0) * (np.arange(T_x)[None, None, :] < np.asarray(t_xs)[:, None, None])).float().unsqueeze(1)
    np.testing.assert_array_equal(oracle.generate_path_torch(w, mask).numpy(), attn.numpy())


@pytest.mark.gpu
def test_generate_path_gpu_matches_golden_and_oracle(oracle):
    import vits_b200
    for name, (dur, mask, path) in _golden().items():
        got = vits_b200.generate_path(torch.from_numpy(dur).cuda(), torch.from_numpy(mask).cuda())
        assert got.shape == path.shape and got.dtype == torch.float32
        np.testing.assert_array_equal(got.cpu().numpy(), path, err_msg=name)
    rng = np.random.default_rng(17)
    for (b, t_x, scale) in [(4, 192, 5.0), (3, 257, 2.0), (2, 33, 9.0)]:
        dur = torch.from_numpy(np.ceil(np.exp(rng.standard_normal((b, 1, t_x))) * scale).astype(np.float32))
        x_len = torch.as_tensor(rng.integers(t_x // 2, t_x + 1, size=b)); x_len[0] = t_x
        x_mask = (torch.arange(t_x)[None, :] < x_len[:, None]).float()[:, None, :]
        dur = dur * x_mask
        y_len = dur.sum((1, 2)).clamp_min(1).long()
        y_mask = (torch.arange(int(y_len.max()))[None, :] < y_len[:, None]).float()[:, None, :]
        attn_mask = x_mask.unsqueeze(2) * y_mask.unsqueeze(-1)            # SynthesizerTrn.py:304, a broadcast view
        want = oracle.generate_path_torch(dur, attn_mask)
        got = vits_b200.generate_path(dur.cuda(), attn_mask.cuda())
        np.testing.assert_array_equal(got.cpu().numpy(), want.numpy())
        # half-precision masks keep their dtype (the reference casts the path to mask.dtype)
        assert vits_b200.generate_path(dur.cuda(), attn_mask.cuda().half()).dtype == torch.float16


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 200, 70, 192), (2, 1024, 192, 192), (2, 333, 257, 5)])
def test_durations_and_prior_expansion_gpu(oracle, shape):
    import vits_b200
    B, T_y, T_x, C = shape
    rng = np.random.default_rng(B * T_y + T_x)
    nc = torch.from_numpy((rng.standard_normal((B, T_y, T_x)) * 3).astype(np.float32)).cuda()
    t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
    ty, tx = torch.as_tensor(t_ys).cuda(), torch.as_tensor(t_xs).cuda()
    path = vits_b200.maximum_path_from_lengths(nc, ty, tx)
    index = vits_b200.maximum_path_index(nc, y_lengths=ty, x_lengths=tx)
    attn = path.unsqueeze(1)
    # durations: bit-exact with attn.sum(2)
    w = vits_b200.path_durations(index, T_x)
    assert w.shape == (B, 1, T_x)
    assert torch.equal(w, oracle.durations_torch(attn))
    # expansion: bit-exact with the einsum on the dense path, for one and for two statistics
    m_p = torch.randn(B, C, T_x, device="cuda", requires_grad=True)
    logs_p = (torch.randn(B, C, T_x, device="cuda") * 0.3).requires_grad_()
    m_e, l_e = vits_b200.expand_prior(index, m_p, logs_p)
    m_ref = oracle.expand_prior_torch(attn, m_p)
    l_ref = oracle.expand_prior_torch(attn, logs_p)
    assert torch.equal(m_e, m_ref.detach()) and torch.equal(l_e, l_ref.detach())
    only, none = vits_b200.expand_prior(index, m_p.detach())
    assert none is None and torch.equal(only, m_ref.detach())
    # gradients: the transpose of the gather == autograd through the einsum (fp32 sums of a few terms)
    g1, g2 = torch.randn_like(m_e), torch.randn_like(l_e)
    (m_e * g1).sum().backward(retain_graph=True)
    (l_e * g2).sum().backward()
    gm, gl = m_p.grad.clone(), logs_p.grad.clone()
    m_p.grad = logs_p.grad = None
    ((m_ref * g1).sum() + (l_ref * g2).sum()).backward()
    torch.testing.assert_close(gm, m_p.grad, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(gl, logs_p.grad, rtol=1e-5, atol=1e-5)


def test_oracle_kl_loss_matches_reference(oracle):
    """losses.py imports only torch: check the restatement against the reference's own function when it is here."""
    import importlib.util
    import sys
    ref = "/root/reference/losses.py"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present (GPU box)")
    spec = importlib.util.spec_from_file_location("ref_losses", ref)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = torch.Generator().manual_seed(5)
    z_p, logs_q, m_p, logs_p = (torch.randn(2, 6, 17, generator=g) * s for s in (1.0, 0.3, 1.0, 0.3))
    z_mask = (torch.arange(17)[None, :] < torch.tensor([17, 11])[:, None]).float()[:, None, :]
    want = mod.kl_loss(z_p, logs_q, m_p, logs_p, z_mask)
    assert torch.equal(oracle.kl_loss_torch(z_p, logs_q, m_p, logs_p, z_mask), want)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 200, 70, 192), (2, 1024, 192, 192), (2, 333, 257, 5)])
def test_kl_loss_from_index_gpu(oracle, shape):
    import vits_b200
    B, T_y, T_x, C = shape
    rng = np.random.default_rng(7 * B + T_x)
    nc = torch.from_numpy((rng.standard_normal((B, T_y, T_x)) * 3).astype(np.float32)).cuda()
    t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
    ty, tx = torch.as_tensor(t_ys).cuda(), torch.as_tensor(t_xs).cuda()
    index = vits_b200.maximum_path_index(nc, y_lengths=ty, x_lengths=tx)
    attn = vits_b200.maximum_path_from_lengths(nc, ty, tx).unsqueeze(1)
    z_mask = (torch.arange(T_y, device="cuda")[None, :] < ty[:, None]).float()[:, None, :]
    leaves = [torch.randn(B, C, T_y, device="cuda").requires_grad_(), (torch.randn(B, C, T_y, device="cuda") * 0.3).requires_grad_(),
              torch.randn(B, C, T_x, device="cuda").requires_grad_(), (torch.randn(B, C, T_x, device="cuda") * 0.3).requires_grad_()]
    z_p, logs_q, m_p, logs_p = leaves
    got = vits_b200.kl_loss_from_index(index, z_p, logs_q, m_p, logs_p, z_mask)
    # the reference: expand with the einsums on the dense path, then losses.kl_loss
    want = oracle.kl_loss_torch(z_p, logs_q, oracle.expand_prior_torch(attn, m_p), oracle.expand_prior_torch(attn, logs_p), z_mask)
    torch.testing.assert_close(got, want.detach(), rtol=2e-5, atol=1e-6)     # summation order differs (fp64 accumulation here)
    got.backward()
    grads = [t.grad.clone() for t in leaves]
    for t in leaves:
        t.grad = None
    want.backward()
    for a, t in zip(grads, leaves):
        torch.testing.assert_close(a, t.grad, rtol=1e-4, atol=1e-7)
