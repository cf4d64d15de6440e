"""GPU parity for the consumers of the alignment (tts.py:503-563): duration
loss, out_size crop, mu_y gather (+ backward), prior loss -- the sm_100a kernels through the
C ABI against the golden capture of the reference's compute_loss and against the CPU oracle.
Bars: integer / gather / crop results bit-exact; fp32 reductions within 1e-5 relative."""
import os
import random

import numpy as np
import pytest
import torch

import oracle
from conftest import GOLDEN, rect_mask

pytestmark = pytest.mark.gpu


def dev_t(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


def test_loss_block_matches_reference_compute_loss(cuda):
    from art_tts_b200 import alignment
    g = np.load(os.path.join(GOLDEN, "loss_block_gradtts.npz"))
    mu_x = dev_t(g["mu_x"], cuda).requires_grad_(True)
    logw = dev_t(g["logw"], cuda).requires_grad_(True)
    y = dev_t(g["y"], cuda)
    x_len, y_len = torch.from_numpy(g["x_lengths"]), torch.from_numpy(g["y_lengths"])
    rng = random.Random(int(g["random_seed"]))   # same stream as the reference's random.seed()
    out = alignment.alignment_losses(mu_x, logw, x_len, y, y_len, int(g["out_size"]),
                                     y_lengths_host=y_len, rng=rng, return_attn=True)
    assert np.array_equal(out.durations.cpu().numpy(), g["durations"])
    assert np.array_equal(out.y.cpu().numpy(), g["y_cut"])
    assert np.array_equal(out.y_mask.cpu().numpy(), g["y_cut_mask"])
    assert np.array_equal(out.mu_y.detach().cpu().numpy(), g["mu_y"])
    assert np.isclose(out.dur_loss.item(), g["dur_loss"], rtol=1e-5)
    assert np.isclose(out.prior_loss.item(), g["prior_loss"], rtol=1e-5)
    # attn_cut @ mu_x reproduces mu_y: the dense crop agrees with the gather
    attn = out.attn.cpu().numpy()
    assert np.array_equal(oracle.align_mu_y(attn, g["mu_x"]), g["mu_y"])
    (out.dur_loss + out.prior_loss).backward()
    torch.cuda.synchronize()
    gm, gl = mu_x.grad.cpu().numpy(), logw.grad.cpu().numpy()
    assert np.allclose(gm, g["grad_mu_x"], rtol=1e-4, atol=1e-8)
    assert np.allclose(gl, g["grad_logw"], rtol=1e-5, atol=1e-8)


def ragged_case(seed, B, F, T_x, T_y):
    rng = np.random.default_rng(seed)
    x_len = rng.integers(max(1, T_x // 4), T_x + 1, B).astype(np.int32)
    y_len = np.minimum(T_y, 3 * x_len + rng.integers(0, 40, B)).astype(np.int32)
    x_len[0], y_len[0] = T_x, T_y
    mu_x = (rng.standard_normal((B, F, T_x)) * (np.arange(T_x)[None, None] < x_len[:, None, None])).astype(np.float32)
    y = (rng.standard_normal((B, F, T_y)) * (np.arange(T_y)[None, None] < y_len[:, None, None])).astype(np.float32)
    return mu_x, y, x_len, y_len


@pytest.mark.parametrize("out_size", [None, 64, 172])
@pytest.mark.parametrize("F,B,T_x,T_y", [(80, 16, 60, 260), (16, 9, 33, 131)])
def test_alignment_consumers_vs_oracle(cuda, F, B, T_x, T_y, out_size):
    from art_tts_b200 import alignment, monotonic_align
    mu_x, y, x_len, y_len = ragged_case(11 + F, B, F, T_x, T_y)
    mu_t, y_t = dev_t(mu_x, cuda), dev_t(y, cuda)
    tx, ty = torch.from_numpy(x_len), torch.from_numpy(y_len)
    path, dur, fidx = monotonic_align.maximum_path_from_prior(mu_t, None, y_t, tx, ty,
                                                              return_frame_idx=True)
    attn = path.cpu().numpy()
    # frame index: fused kernel output == oracle's == the durations -> index kernel
    want_idx = oracle.frame_index(attn, y_len)
    assert np.array_equal(fidx.cpu().numpy(), want_idx)
    assert np.array_equal(alignment.frame_index(dur, tx, ty, T_y).cpu().numpy(), want_idx)
    # duration targets / loss
    x_mask = (np.arange(T_x)[None, None] < x_len[:, None, None]).astype(np.float32)
    logw = np.random.default_rng(3).standard_normal((B, 1, T_x)).astype(np.float32) * x_mask
    logw_ = oracle.duration_targets(attn, x_mask)
    got = alignment.duration_targets(dur, tx).cpu().numpy()
    assert np.allclose(got, logw_, rtol=2e-7, atol=1e-7)
    lw = dev_t(logw, cuda).requires_grad_(True)
    dl = alignment.duration_loss_from_durations(lw, dur, tx)
    assert np.isclose(dl.item(), oracle.duration_loss(logw, logw_, x_len), rtol=1e-5)
    dl.backward()
    assert np.allclose(lw.grad.cpu().numpy(), 2 * (logw - logw_) / x_len.sum(), rtol=1e-5, atol=1e-8)
    # crop
    if out_size is None:
        T_out, off, seg = T_y, None, y_len.astype(np.int64)
        y_cut, attn_cut = y, attn
        off_t = None
    else:
        T_out = out_size
        off = oracle.crop_offsets(y_len, out_size, random.Random(5))
        offs2, lens2 = alignment.crop_offsets(torch.from_numpy(y_len), out_size, random.Random(5))
        y_cut, attn_cut, seg = oracle.crop_segments(y, attn, y_len, out_size, off)
        assert offs2 == off.tolist() and lens2 == seg.tolist()
        off_t = torch.tensor(offs2, dtype=torch.int32)
        assert np.array_equal(alignment.crop(y_t, off_t, torch.tensor(lens2), out_size).cpu().numpy(), y_cut)
    seg_t = torch.from_numpy(np.asarray(seg, np.int32))
    got_attn = alignment.path_segment(fidx, off_t, seg_t, T_x, T_out).cpu().numpy()
    assert np.array_equal(got_attn, attn_cut)
    for dt in (torch.uint8, torch.float16, torch.int32):
        assert np.array_equal(alignment.path_segment(fidx, off_t, seg_t, T_x, T_out, dt).cpu().numpy(),
                              attn_cut.astype(got_attn.dtype))
    # mu_y gather == one-hot GEMM, prior loss, and both gradients
    mu_g = dev_t(mu_x, cuda).requires_grad_(True)
    mu_y, pl = alignment.aligned_mu_y_and_prior_loss(mu_g, dev_t(y_cut, cuda), fidx, off_t, seg_t, T_out)
    want_mu_y = oracle.align_mu_y(attn_cut, mu_x)
    assert np.array_equal(mu_y.detach().cpu().numpy(), want_mu_y)
    y_mask = oracle.sequence_mask(seg, T_out)[:, None, :].astype(np.float32)
    assert np.isclose(pl.item(), oracle.prior_loss(y_cut, want_mu_y, y_mask, F), rtol=1e-5)
    g_dec = np.random.default_rng(4).standard_normal(want_mu_y.shape).astype(np.float32)
    (pl * 3.0 + (mu_y * dev_t(g_dec, cuda)).sum()).backward()
    g_mu_y = g_dec.astype(np.float64) + 3.0 * (want_mu_y - y_cut) * y_mask / (y_mask.sum() * F)
    want_grad = oracle.align_mu_y_grad(attn_cut, g_mu_y)
    assert np.allclose(mu_g.grad.cpu().numpy(), want_grad, rtol=1e-4, atol=1e-5)
    # gather only (no loss): same values, gradient = segmented sum of the incoming gradient
    mu_g2 = dev_t(mu_x, cuda).requires_grad_(True)
    mu_y2 = alignment.aligned_mu_y(mu_g2, fidx, off_t, seg_t, T_out)
    assert np.array_equal(mu_y2.detach().cpu().numpy(), want_mu_y)
    (mu_y2 * dev_t(g_dec, cuda)).sum().backward()
    assert np.allclose(mu_g2.grad.cpu().numpy(), oracle.align_mu_y_grad(attn_cut, g_dec), rtol=1e-4, atol=1e-5)


def test_alignment_block_at_b1024_properties(cuda):
    """Size-independent checks at the bench shape: every valid frame gathers its own token's
    mu_x, padding is zero, and prior_loss equals the masked mean of -(log-prior of the path)/F."""
    from art_tts_b200 import alignment
    B, F, T_x, T_y = 1024, 80, 190, 872
    rng = np.random.default_rng(0)
    x_len = rng.integers(60, T_x + 1, B).astype(np.int32)
    y_len = np.minimum(T_y, 4 * x_len + rng.integers(0, 100, B)).astype(np.int32)
    tx, ty = torch.from_numpy(x_len).to(cuda), torch.from_numpy(y_len).to(cuda)
    gen = torch.Generator(device=cuda).manual_seed(1)
    mu_x = torch.randn(B, F, T_x, device=cuda, generator=gen)
    y = torch.randn(B, F, T_y, device=cuda, generator=gen)
    logw = torch.zeros(B, 1, T_x, device=cuda)
    out = alignment.alignment_losses(mu_x, logw, tx, y, ty)
    valid = torch.arange(T_y, device=cuda)[None, :] < ty[:, None]
    idx = out.frame_idx.long().clamp(min=0)
    want = torch.gather(mu_x, 2, idx[:, None, :].expand(B, F, T_y)) * valid[:, None, :]
    assert torch.equal(out.mu_y, want)
    ref = (0.5 * ((y - want) ** 2 + float(np.log(2 * np.pi))) * valid[:, None, :]).double().sum() / (valid.sum() * F)
    assert abs(out.prior_loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert torch.equal(out.durations.sum(1), ty)


def test_inference_alignment_matches_reference_forward(cuda):
    """tts.py:123-153 (durations -> y_lengths -> generate_path -> mu_y) against captures of the reference's own
    ArtTTS.forward with the decoder replaced by the identity (tests/golden/inference_arttts.npz): given and
    predicted durations, length_scale 1, 1.3 and 2."""
    from art_tts_b200 import alignment
    g = np.load(os.path.join(GOLDEN, "inference_arttts.npz"))
    mu_x = torch.from_numpy(g["mu_x"]).to(cuda)
    logw = torch.from_numpy(g["logw"]).to(cuda)
    x_mask = torch.from_numpy(g["x_mask"]).to(cuda)
    durs = torch.from_numpy(g["x_durations"]).to(cuda)
    T_x = mu_x.shape[2]
    for name in g["names"]:
        given = str(name).startswith("given")
        ls = float(g[f"{name}.length_scale"])
        out = alignment.inference_alignment(mu_x, logw, x_mask, length_scale=ls, x_durations=durs if given else None)
        T_y = int(g[f"{name}.T_y"])
        want_attn = np.unpackbits(g[f"{name}.attn_packed"], axis=-1)[:, :, :T_y].astype(np.float32)
        assert out.attn.shape == (mu_x.shape[0], T_x, T_y), name
        assert np.array_equal(out.attn.cpu().numpy(), want_attn), name
        want_mu_y = g[f"{name}.mu_y"]                       # [B, F, y_max_length]
        assert out.y_max_length == want_mu_y.shape[2], name
        assert np.array_equal(out.mu_y[:, :, :out.y_max_length].cpu().numpy(), want_mu_y), name
        assert float(out.mu_y[:, :, out.y_max_length:].abs().sum()) == 0.0, name
