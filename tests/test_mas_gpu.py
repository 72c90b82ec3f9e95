"""GPU parity tests of maximum_path: the CUDA path (through the drop-in wrapper and the raw C ABI)
against the CPU oracle on identical fp32 inputs.  Bar: bit-exact paths."""
import ctypes

import numpy as np
import pytest
import torch

from helpers import check_path_properties, index_to_path, path_to_index, random_lengths, sha_path

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mp():
    import vits_b200.monotonic_align as m
    yield m
    m._lib.lib().mas_set_tuning(0, 0, 0, -1)
    m._lib.lib().mas_set_tuning2(-1, 0)


def _mask(t_ys, t_xs, T_y, T_x, device, dtype=torch.float32):
    ty = torch.as_tensor(t_ys, device=device)
    tx = torch.as_tensor(t_xs, device=device)
    ym = torch.arange(T_y, device=device)[None, :] < ty[:, None]
    xm = torch.arange(T_x, device=device)[None, :] < tx[:, None]
    return (ym[:, :, None] & xm[:, None, :]).to(dtype)


def _gpu_path(mp, nc_np, t_ys, t_xs, via="mask"):
    nc = torch.from_numpy(nc_np).cuda()
    B, T_y, T_x = nc.shape
    if via == "mask":
        out = mp.maximum_path(nc, _mask(t_ys, t_xs, T_y, T_x, nc.device))
    else:
        out = mp.maximum_path_from_lengths(nc, torch.as_tensor(t_ys), torch.as_tensor(t_xs))
    assert out.dtype == nc.dtype and out.device == nc.device and out.shape == nc.shape
    return out.cpu().numpy().astype(np.int8)


def test_golden_vectors(mp, mas_golden):
    for name, c in mas_golden.items():
        for via in ("mask", "lengths"):
            got = _gpu_path(mp, c["neg_cent"], c["t_ys"], c["t_xs"], via)
            assert sha_path(got) == c["sha256"], (name, via)


SWEEP = [(3, 9, 1), (2, 40, 2), (4, 64, 31), (4, 64, 32), (3, 70, 33), (2, 130, 63), (2, 130, 64), (2, 131, 65),
         (2, 97, 96), (2, 97, 97), (2, 98, 97), (2, 400, 191), (2, 400, 192), (2, 401, 193), (2, 700, 256),
         (2, 777, 257), (1, 1030, 449), (2, 1100, 512), (1, 1200, 900), (1, 2100, 1024)]


@pytest.mark.parametrize("shape", SWEEP)
def test_random_sweep_bit_exact(mp, oracle, shape):
    B, T_y, T_x = shape
    rng = np.random.default_rng(hash(shape) % (2 ** 32))
    nc = (rng.standard_normal(shape) * 3 - 4).astype(np.float32)
    t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
    want = oracle.maximum_path_numpy(nc, t_ys, t_xs).astype(np.int8)
    got = _gpu_path(mp, nc, t_ys, t_xs)
    np.testing.assert_array_equal(got, want)
    # minimal and degenerate lengths: t_x == t_y (pure diagonal), t_y == t_x + 1, t_x == 1
    t_xs2 = np.minimum(t_xs, T_y)
    t_ys2 = t_xs2.copy()
    if B > 1:
        t_ys2[1] = min(T_y, t_xs2[1] + 1)
    if B > 2:
        t_xs2[2] = 1
    want = oracle.maximum_path_numpy(nc, t_ys2, t_xs2).astype(np.int8)
    got = _gpu_path(mp, nc, t_ys2, t_xs2, via="lengths")
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("K", [1, 2, 3, 4, 6, 8])
@pytest.mark.parametrize("R", [8, 16, 32])
def test_every_kernel_configuration(mp, oracle, K, R):
    """Each cols-per-lane / rows-per-stage instantiation, vector and scalar load paths."""
    L = mp._lib.lib()
    rng = np.random.default_rng(1000 * K + R)
    for shape in [(3, 150, 96), (2, 333, 100), (2, 260, 193)]:
        B, T_y, T_x = shape
        nc = (rng.standard_normal(shape) * 2 - 1).astype(np.float32)
        t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
        want = oracle.maximum_path_numpy(nc, t_ys, t_xs).astype(np.int8)
        for stages, pdl, fused, helpers in ((2, 1, 0, 0), (5, 0, 0, 0), (3, 1, 1, 1), (4, 0, 1, 4), (0, 1, -1, 0),
                                            (3, 1, 2, 0), (0, 0, 2, 0), (0, 1, 3, 0), (0, 2, -1, 0), (3, 2, 1, 1)):
            L.mas_set_tuning(K, R, stages, pdl)
            L.mas_set_tuning2(fused, helpers)
            got = _gpu_path(mp, nc, t_ys, t_xs)
            np.testing.assert_array_equal(got, want, err_msg=f"K={K} R={R} S={stages} pdl={pdl} fused={fused} {shape}")
    L.mas_set_tuning(0, 0, 0, -1)
    L.mas_set_tuning2(-1, 0)


@pytest.mark.parametrize("ring_mode", [1, 2, 3, 4])
@pytest.mark.parametrize("K", [1, 2, 4])
def test_wavefront_forward_kernel_configurations(mp, oracle, K, ring_mode):
    """The wavefront forward kernel (mode 3): every columns-per-lane instantiation, the linear ring at skew 1..3
    and the select ring, minimal and deep rings, TMA and cp.async chunk loads (T_x % 4 != 0), with/without PDL."""
    L = mp._lib.lib()
    rng = np.random.default_rng(50 * K + ring_mode)
    smin = {1: 3, 2: 4, 3: 5, 4: 3}[ring_mode]
    try:
        for shape in [(3, 150, 96), (2, 333, 100), (2, 260, 129), (2, 300, 31), (2, 700, 190)]:
            B, T_y, T_x = shape
            nc = (rng.standard_normal(shape) * 2 - 1).astype(np.float32)
            t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
            want = oracle.maximum_path_numpy(nc, t_ys, t_xs).astype(np.int8)
            for slots, pdl in ((smin, 1), (0, 0), (0, 1), (0, 2)):   # pdl 2: forward kernel launched programmatically too
                L.mas_set_tuning(0, 0, 0, pdl)
                L.mas_set_tuning2(3, 0)
                L.mas_set_tuning3(-1, ring_mode, slots, K)
                got = _gpu_path(mp, nc, t_ys, t_xs)
                np.testing.assert_array_equal(got, want, err_msg=f"K={K} ring={ring_mode} S={slots} pdl={pdl} {shape}")
    finally:
        L.mas_set_tuning(0, 0, 0, -1)
        L.mas_set_tuning2(-1, 0)
        L.mas_set_tuning3(-1, 0, 0, 0)


@pytest.mark.parametrize("wavefront", [1, 32, 33, 41, 49])
def test_wavefront_kernel_generations(mp, oracle, wavefront):
    """Both generations of the wavefront forward kernel (mas_set_tuning3: 1 = mas_dp_kernel, 32 = mas_dp2_kernel, 33 =
    mas_dp2_kernel with the instruction-cache warmer = the automatic choice: one CTA per utterance up to four DP warps,
    a cluster of two beyond; 41 = never a cluster, 49 = a cluster whenever the text has two warps) against the oracle:
    ordinary values, ties, values below the -1e9 sentinel (core.pyx:17-27 -- the diagonal and x == 0 rules are exercised
    exactly there), with the lengths given and taken from the mask, skew 1 and 2."""
    L = mp._lib.lib()
    rng = np.random.default_rng(700 + wavefront)
    try:
        for shape in [(3, 200, 64), (2, 333, 100), (2, 900, 192), (2, 450, 256), (2, 40, 33), (1, 5, 4), (5, 1024, 192),
                      (2, 700, 300), (2, 600, 512), (3, 520, 449)]:
            B, T_y, T_x = shape
            for kind in ("normal", "ties", "subsentinel"):
                if kind == "normal":
                    nc = (rng.standard_normal(shape) * 3 - 4).astype(np.float32)
                elif kind == "ties":
                    nc = rng.integers(-3, 1, size=shape).astype(np.float32)
                else:
                    nc = (rng.standard_normal(shape) * 1e8 - 2e9).astype(np.float32)
                t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
                want = oracle.maximum_path_numpy(nc, t_ys, t_xs).astype(np.int8)
                for skew in (0, 1, 2):
                    L.mas_set_tuning3(wavefront, skew, 0, 0)
                    for via in ("mask", "lengths"):
                        got = _gpu_path(mp, nc, t_ys, t_xs, via=via)
                        np.testing.assert_array_equal(got, want, err_msg=f"wavefront={wavefront} skew={skew} {kind} {via} {shape}")
    finally:
        L.mas_set_tuning3(-1, 0, 0, 0)


def test_streaming_backtrack_with_16_bit_tables(mp, oracle):
    """Long utterances: the streaming backtrack keeps 16-bit exit columns and re-walks the groups."""
    L = mp._lib.lib()
    rng = np.random.default_rng(4096)
    # (2, 5000, 300): every group's 32-word walk window fits in shared memory next to the tables; (1, 4800, 512): it does
    # not, the groups are re-walked from the scratch
    for B, T_y, T_x in ((2, 5000, 300), (1, 4800, 512)):
        nc = (rng.standard_normal((B, T_y, T_x)) * 3).astype(np.float32)
        t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
        want = oracle.maximum_path_numpy(nc, t_ys, t_xs)
        try:
            for mode in (2, -1, 0):
                L.mas_set_tuning2(mode, 0)
                ncd = torch.from_numpy(nc).cuda()
                got = mp.maximum_path_from_lengths(ncd, torch.as_tensor(t_ys), torch.as_tensor(t_xs))
                np.testing.assert_array_equal(got.cpu().numpy().astype(np.int32), want, err_msg=f"mode {mode} {T_y}x{T_x}")
                idx = mp.maximum_path_index(ncd, y_lengths=torch.as_tensor(t_ys), x_lengths=torch.as_tensor(t_xs))
                np.testing.assert_array_equal(idx.cpu().numpy(), path_to_index(want), err_msg=f"mode {mode} {T_y}x{T_x}")
        finally:
            L.mas_set_tuning2(-1, 0)


def test_more_utterances_than_sms(mp, oracle):
    """B >= SM count: the forward CTAs run in waves, so the ones are dropped by the write-out kernel, not by
    the forward kernel (which would wait for a zero-fill that cannot start)."""
    rng = np.random.default_rng(123)
    B, T_y, T_x = 333, 96, 40
    nc = (rng.standard_normal((B, T_y, T_x)) * 2).astype(np.float32)
    t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
    want = oracle.maximum_path_numpy(nc, t_ys, t_xs).astype(np.int8)
    for _ in range(2):
        np.testing.assert_array_equal(_gpu_path(mp, nc, t_ys, t_xs), want)


def test_wide_texts_on_clusters_in_waves_and_short_utterances(mp, oracle):
    """Texts of more than four DP warps run on a cluster of two CTAs per utterance (mas_dp2.cuh): more clusters than the
    machine holds at once (waves), and utterances far shorter than the padded text (t_x <= t_y << T_x: the second CTA's
    columns are all padding, the first chunks lie above the diagonal)."""
    rng = np.random.default_rng(321)
    B, T_y, T_x = 100, 330, 320
    nc = (rng.standard_normal((B, T_y, T_x)) * 2 - 1).astype(np.float32)
    t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
    want = oracle.maximum_path_numpy(nc, t_ys, t_xs).astype(np.int8)
    for via in ("mask", "lengths"):
        np.testing.assert_array_equal(_gpu_path(mp, nc, t_ys, t_xs, via=via), want)
    B, T_y, T_x = 6, 90, 400
    nc = (rng.standard_normal((B, T_y, T_x)) * 2 - 1).astype(np.float32)
    t_ys = np.array([90, 70, 64, 33, 5, 1], np.int32)
    t_xs = np.array([90, 40, 64, 32, 3, 1], np.int32)
    want = oracle.maximum_path_numpy(nc, t_ys, t_xs).astype(np.int8)
    for via in ("mask", "lengths"):
        np.testing.assert_array_equal(_gpu_path(mp, nc, t_ys, t_xs, via=via), want)


def test_back_to_back_calls_share_scratch(mp, oracle):
    """Consecutive calls overlap through programmatic dependent launch and reuse one scratch buffer:
    enqueue many without synchronising, with and without the dense path, then check every result."""
    rng = np.random.default_rng(77)
    B, T_y, T_x = 6, 420, 160
    ncs, lens, wants = [], [], []
    for _ in range(12):
        nc = (rng.standard_normal((B, T_y, T_x)) * 3).astype(np.float32)
        t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
        ncs.append(torch.from_numpy(nc).cuda())
        lens.append((torch.as_tensor(t_ys).cuda(), torch.as_tensor(t_xs).cuda()))
        wants.append(oracle.maximum_path_numpy(nc, t_ys, t_xs))
    torch.cuda.synchronize()
    for rep in range(3):
        outs = []
        for i, nc in enumerate(ncs):
            if (i + rep) % 3 == 2:
                outs.append(("index", mp.maximum_path_index(nc, y_lengths=lens[i][0], x_lengths=lens[i][1])))
            else:
                outs.append(("path", mp.maximum_path_from_lengths(nc, lens[i][0], lens[i][1])))
        torch.cuda.synchronize()
        for (kind, out), want in zip(outs, wants):
            if kind == "path":
                np.testing.assert_array_equal(out.cpu().numpy().astype(np.int32), want)
            else:
                np.testing.assert_array_equal(out.cpu().numpy(), path_to_index(want))


def test_ties_and_huge_magnitudes(mp, oracle):
    rng = np.random.default_rng(11)
    nc = rng.integers(-2, 3, size=(4, 300, 150)).astype(np.float32)          # dense ties
    t_ys, t_xs = random_lengths(rng, 4, 300, 150)
    np.testing.assert_array_equal(_gpu_path(mp, nc, t_ys, t_xs), oracle.maximum_path_numpy(nc, t_ys, t_xs).astype(np.int8))
    nc = np.zeros((2, 200, 70), np.float32)                                  # all ties
    np.testing.assert_array_equal(_gpu_path(mp, nc, [200, 150], [70, 70]),
                                  oracle.maximum_path_numpy(nc, [200, 150], [70, 70]).astype(np.int8))
    nc = (rng.standard_normal((3, 250, 90)) * 5e8 - 1e9).astype(np.float32)  # values below the -1e9 sentinel
    t_ys, t_xs = random_lengths(rng, 3, 250, 90)
    np.testing.assert_array_equal(_gpu_path(mp, nc, t_ys, t_xs), oracle.maximum_path_numpy(nc, t_ys, t_xs).astype(np.int8))


def test_dtypes_and_mask_kinds(mp, oracle):
    rng = np.random.default_rng(5)
    B, T_y, T_x = 3, 120, 48
    nc32 = torch.from_numpy((rng.standard_normal((B, T_y, T_x)) * 2).astype(np.float32)).cuda()
    t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
    for dt in (torch.float32, torch.float16, torch.bfloat16, torch.float64):
        nc = nc32.to(dt)
        keep = nc.clone()
        want = oracle.maximum_path_numpy(nc.float().cpu().numpy(), t_ys, t_xs)
        for mdt in (torch.float32, torch.bool, torch.float16, torch.int64, torch.uint8):
            out = mp.maximum_path(nc, _mask(t_ys, t_xs, T_y, T_x, nc.device, mdt))
            assert out.dtype == dt and out.device == nc.device
            np.testing.assert_array_equal(out.cpu().double().numpy(), want.astype(np.float64))
        assert torch.equal(nc, keep), "input must not be modified"
    # the caller's mask is a broadcast product of two prefix masks (SynthesizerTrn.py:234); an expanded
    # (stride-0) view must work too
    xm = (torch.arange(T_x, device="cuda")[None, :] < torch.as_tensor(t_xs, device="cuda")[:, None]).float()
    ym = (torch.arange(T_y, device="cuda")[None, :] < torch.as_tensor(t_ys, device="cuda")[:, None]).float()
    attn_mask = (xm[:, None, None, :] * ym[:, None, :, None]).squeeze(1)
    want = oracle.maximum_path_numpy(nc32.cpu().numpy(), t_ys, t_xs)
    np.testing.assert_array_equal(mp.maximum_path(nc32, attn_mask).cpu().numpy(), want.astype(np.float32))
    col = ym[:, :, None].expand(B, T_y, T_x)  # stride 0 along x: column 0 still carries t_y ...
    row = xm[:, None, :].expand(B, T_y, T_x)
    np.testing.assert_array_equal(mp.maximum_path(nc32, col * row).cpu().numpy(), want.astype(np.float32))


def test_misaligned_base_pointer(mp, oracle):
    """A contiguous view whose storage offset breaks 16-byte alignment takes the scalar-load path."""
    rng = np.random.default_rng(8)
    B, T_y, T_x = 2, 200, 64
    for off in (1, 2, 3):
        flat = torch.from_numpy(rng.standard_normal(B * T_y * T_x + off).astype(np.float32)).cuda()
        nc = flat[off:].view(B, T_y, T_x)
        assert nc.is_contiguous() and nc.data_ptr() % 16 != 0
        t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
        want = oracle.maximum_path_numpy(nc.cpu().numpy(), t_ys, t_xs).astype(np.int8)
        got = mp.maximum_path_from_lengths(nc, torch.as_tensor(t_ys), torch.as_tensor(t_xs)).cpu().numpy().astype(np.int8)
        np.testing.assert_array_equal(got, want)


def test_index_output(mp, oracle):
    rng = np.random.default_rng(21)
    B, T_y, T_x = 4, 310, 130
    nc = (rng.standard_normal((B, T_y, T_x))).astype(np.float32)
    t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
    idx = mp.maximum_path_index(torch.from_numpy(nc).cuda(), y_lengths=torch.as_tensor(t_ys), x_lengths=torch.as_tensor(t_xs))
    assert idx.dtype == torch.int32 and tuple(idx.shape) == (B, T_y)
    want = oracle.maximum_path_numpy(nc, t_ys, t_xs)
    np.testing.assert_array_equal(idx.cpu().numpy(), path_to_index(want))


def test_invalid_lengths_flagged_not_imitated(mp):
    """t_x > t_y and empty utterances are UB in the reference: we zero the path and flag."""
    nc = torch.randn(3, 20, 12, device="cuda")
    t_ys = torch.tensor([20, 5, 0])
    t_xs = torch.tensor([12, 9, 4])
    mp.last_status()
    out = mp.maximum_path_from_lengths(nc, t_ys, t_xs)
    assert out[0].sum().item() == 20 and out[1].sum().item() == 0 and out[2].sum().item() == 0
    bits = mp.last_status()
    assert bits & 1 and bits & 2
    assert mp.last_status() == 0
    with pytest.raises(mp.StatusError):
        mp.maximum_path_from_lengths(nc, t_ys, t_xs, check=True)
    assert mp.maximum_path_from_lengths(nc, torch.tensor([20, 20, 20]), torch.tensor([12, 9, 4]), check=True).sum().item() == 60


def test_invalid_lengths_on_a_cluster_sized_text(mp, oracle):
    """The same, where an utterance runs on a cluster of two CTAs with a courier warp at the boundary (mas_dp2.cuh): empty
    and t_x > t_y utterances next to valid ones, lengths given and taken from the mask -- every warp of both CTAs must
    find its way out, and the valid utterances keep their exact paths."""
    rng = np.random.default_rng(99)
    B, T_y, T_x = 6, 500, 400
    nc = (rng.standard_normal((B, T_y, T_x)) * 2).astype(np.float32)
    t_ys = np.array([500, 0, 120, 33, 400, 1], np.int32)
    t_xs = np.array([400, 300, 200, 33, 399, 1], np.int32)       # utterance 2: t_x > t_y; utterance 1: empty
    ok = [0, 3, 4, 5]
    want = oracle.maximum_path_numpy(nc[ok], t_ys[ok], t_xs[ok]).astype(np.int8)
    ncd = torch.from_numpy(nc).cuda()
    ty, tx = torch.as_tensor(t_ys).cuda(), torch.as_tensor(t_xs).cuda()
    mask = ((torch.arange(T_y, device="cuda")[None, :] < ty[:, None])[:, :, None]
            & (torch.arange(T_x, device="cuda")[None, :] < tx[:, None])[:, None, :]).float()
    for rep in range(3):
        mp.last_status()
        for out in (mp.maximum_path_from_lengths(ncd, ty, tx), mp.maximum_path(ncd, mask)):
            got = out.cpu().numpy().astype(np.int8)
            np.testing.assert_array_equal(got[ok], want)
            assert got[1].sum() == 0 and got[2].sum() == 0
        bits = mp.last_status()
        assert bits & 1 and bits & 2 and not bits & 8
    assert (mp.status_nosync(reset=True) & 3) == 3   # (the device-wide mirror saw them too; left clean for the next test)


def test_raw_c_abi_overwrites_output_and_respects_stream(mp, oracle):
    L = mp._lib.lib()
    rng = np.random.default_rng(31)
    B, T_y, T_x = 3, 257, 100
    nc_np = rng.standard_normal((B, T_y, T_x)).astype(np.float32)
    t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
    want = oracle.maximum_path_numpy(nc_np, t_ys, t_xs)
    nc = torch.from_numpy(nc_np).cuda()
    ty = torch.as_tensor(t_ys, dtype=torch.int32).cuda()
    tx = torch.as_tensor(t_xs, dtype=torch.int32).cuda()
    scratch = torch.zeros(L.mas_maximum_path_scratch_bytes(B, T_y, T_x), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()
    for dt, code in ((torch.int32, 7), (torch.int8, 5), (torch.float32, 0), (torch.float64, 3), (torch.int64, 8), (torch.float16, 1)):
        path = torch.full((B, T_y, T_x), 77, dtype=dt, device="cuda")          # poison: must be fully overwritten
        idx = torch.full((B, T_y), -7, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            rc = L.mas_maximum_path(nc.data_ptr(), ty.data_ptr(), tx.data_ptr(), None, 0, 0, 0, 0, path.data_ptr(), code,
                                    idx.data_ptr(), scratch.data_ptr(), scratch.numel(), B, T_y, T_x, stream.cuda_stream)
        assert rc == 0
        stream.synchronize()
        np.testing.assert_array_equal(path.cpu().numpy().astype(np.int32), want)
        np.testing.assert_array_equal(idx.cpu().numpy(), path_to_index(want))
    assert L.mas_launch_count() > 0


def test_host_buffer_entry_matches_maximum_path_c(mp, oracle):
    """mas_maximum_path_c_host is the twin of core.pyx:38 on host arrays; the wrapper uses it for CPU tensors."""
    L = mp._lib.lib()
    rng = np.random.default_rng(41)
    for (B, T_y, T_x) in [(1, 50, 20), (5, 180, 77), (19, 300, 128)]:
        values = rng.standard_normal((B, T_y, T_x)).astype(np.float32)
        keep = values.copy()
        t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
        paths = np.zeros((B, T_y, T_x), dtype=np.int32)  # the caller zeroes it, __init__.py:15
        rc = L.mas_maximum_path_c_host(paths.ctypes.data, values.ctypes.data, t_ys.ctypes.data, t_xs.ctypes.data, B, T_y, T_x)
        assert rc == 0
        want = oracle.maximum_path_numpy(keep, t_ys, t_xs)
        np.testing.assert_array_equal(paths, want)
        np.testing.assert_array_equal(values, keep)
        # rows below t_y are written in full; rows at or beyond it are either left alone (like core.pyx:13-33)
        # or zeroed, and nothing beyond the longest utterance is touched
        dirty = np.full((B, T_y, T_x), 9, dtype=np.int32)
        assert L.mas_maximum_path_c_host(dirty.ctypes.data, values.ctypes.data, t_ys.ctypes.data, t_xs.ctypes.data, B, T_y, T_x) == 0
        for b in range(B):
            np.testing.assert_array_equal(dirty[b, :t_ys[b]], want[b, :t_ys[b]])
            assert np.isin(dirty[b, t_ys[b]:], (0, 9)).all()
            assert (dirty[b, t_ys.max():] == 9).all()
        nc_cpu = torch.from_numpy(values)
        out = mp.maximum_path(nc_cpu, _mask(t_ys, t_xs, T_y, T_x, "cpu"))
        assert out.device.type == "cpu" and out.dtype == torch.float32
        np.testing.assert_array_equal(out.numpy().astype(np.int32), paths)
    L.mas_host_release()


CONFIGS = {"c1": (1, 128, 32), "c2": (64, 1024, 192), "c3": (32, 1536, 256), "c4": (8, 4096, 512)}


@pytest.mark.parametrize("cfg", list(CONFIGS))
@pytest.mark.parametrize("ragged", [False, True])
def test_baseline_configs_full_size(mp, oracle, cfg, ragged):
    """BASELINE.json configs at full size: bit-exact vs the oracle AND the size-independent invariants."""
    B, T_y, T_x = CONFIGS[cfg]
    g = torch.Generator(device="cuda").manual_seed(1234)
    nc = torch.randn(B, T_y, T_x, generator=g, device="cuda") * 20 - 400
    rng = np.random.default_rng(7)
    if ragged:
        t_ys, t_xs = random_lengths(rng, B, T_y, T_x)
    else:
        t_ys, t_xs = np.full(B, T_y, np.int32), np.full(B, T_x, np.int32)
    out = mp.maximum_path(nc, _mask(t_ys, t_xs, T_y, T_x, nc.device))
    got = out.cpu().numpy().astype(np.int8)
    check_path_properties(got, t_ys, t_xs)
    want = oracle.maximum_path_numpy(nc.cpu().numpy(), t_ys, t_xs).astype(np.int8)
    np.testing.assert_array_equal(got, want)
    # idempotence / determinism: a second call gives the same bytes
    again = mp.maximum_path(nc, _mask(t_ys, t_xs, T_y, T_x, nc.device))
    assert torch.equal(out, again)


def test_watchdog_timeout_is_loud_and_never_yields_a_garbage_path(oracle):
    """A streaming-backtrack poll that gives up (here: the forward kernel is withheld through the test hook) must leave an
    all-zero path / index -1, raise MAS_STATUS_TIMEOUT in the scratch word and in the host-mapped mirror, and make the NEXT
    call fail loudly without any synchronisation -- never a path walked over words that did not arrive."""
    import vits_b200
    from vits_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(5)
    B, T_y, T_x = 2, 160, 40
    nc = torch.from_numpy((rng.standard_normal((B, T_y, T_x)) * 3).astype(np.float32)).cuda()
    t_ys, t_xs = torch.tensor([160, 121]), torch.tensor([40, 33])
    want = oracle.maximum_path_numpy(nc.cpu().numpy(), t_ys.numpy(), t_xs.numpy())
    vits_b200.status_nosync(reset=True)
    side = torch.cuda.Stream()                  # fresh stream = fresh, zeroed scratch: no stale tags to accept
    L.mas_set_debug_kernels(7 | 8)
    try:
        with torch.cuda.stream(side):
            path = vits_b200.maximum_path_from_lengths(nc, t_ys, t_xs)
            side.synchronize()
            assert int(path.abs().sum().item()) == 0
            assert vits_b200.last_status(reset=True) & _lib.MAS_STATUS_TIMEOUT
    finally:
        L.mas_set_debug_kernels(7)
    assert vits_b200.status_nosync() & _lib.MAS_STATUS_TIMEOUT
    with pytest.raises(_lib.MasError):
        vits_b200.maximum_path_from_lengths(nc, t_ys, t_xs)      # surfaced on the next call, mirror cleared
    got = vits_b200.maximum_path_from_lengths(nc, t_ys, t_xs)
    np.testing.assert_array_equal(got.cpu().numpy().astype(np.int32), want)
    assert vits_b200.status_nosync(reset=True) & _lib.MAS_STATUS_TIMEOUT == 0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process(oracle):
    """Per-device host state (shared-memory opt-ins, SM counts, host-entry buffers) -- ADVICE r1."""
    import vits_b200
    rng = np.random.default_rng(9)
    B, T_y, T_x = 3, 700, 192
    nc = (rng.standard_normal((B, T_y, T_x)) * 3).astype(np.float32)
    t_ys, t_xs = np.array([700, 650, 400], np.int32), np.array([192, 100, 150], np.int32)
    want = oracle.maximum_path_numpy(nc, t_ys, t_xs)
    for d in (1, 0, 1):
        dev = torch.device("cuda", d)
        got = vits_b200.maximum_path_from_lengths(torch.from_numpy(nc).to(dev), torch.as_tensor(t_ys), torch.as_tensor(t_xs))
        assert got.device == dev
        np.testing.assert_array_equal(got.cpu().numpy().astype(np.int32), want)
        z = torch.randn(2, 192, 300, device=dev)
        m, ls = torch.randn(2, 192, 64, device=dev), torch.randn(2, 192, 64, device=dev) * 0.3
        ref = oracle.neg_cent_torch(z, m, ls)
        assert ((vits_b200.neg_cent(z, m, ls) - ref).abs().amax() / ref.abs().amax()).item() <= 1e-5
        with torch.cuda.device(dev):
            cpu = vits_b200.maximum_path_from_lengths(torch.from_numpy(nc), torch.as_tensor(t_ys), torch.as_tensor(t_xs))
        np.testing.assert_array_equal(cpu.numpy().astype(np.int32), want)
