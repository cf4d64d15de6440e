"""GPU parity: the sm_100a kernels (through the C ABI) against the CPU oracle and the golden
vectors recorded from the reference.  Bars (BASELINE.json north_star):
  maximum_path on the reference's own value tensor .... bit-exact path and durations
  fused prior+MAS ..... prior <= 1e-5 relative, >= 99.9 % of cells equal, score <= 1e-4 rel."""
import hashlib
import os

import numpy as np
import pytest
import torch

import oracle
from conftest import GOLDEN, cfg3_inputs, long_text_inputs, path_from_durations, rect_mask, seeded_case

pytestmark = pytest.mark.gpu


def run_mas(value, mask, cuda, **kw):
    from art_tts_b200 import monotonic_align
    v = torch.from_numpy(np.ascontiguousarray(value)).to(cuda)
    m = torch.from_numpy(np.ascontiguousarray(mask)).to(cuda)
    out = monotonic_align.maximum_path(v, m, **kw)
    torch.cuda.synchronize()
    return out


def run_lengths(value, t_x, t_y, cuda, **kw):
    from art_tts_b200 import monotonic_align
    v = torch.from_numpy(np.ascontiguousarray(value)).to(cuda)
    out = monotonic_align.maximum_path_lengths(v, torch.from_numpy(t_x), torch.from_numpy(t_y), **kw)
    torch.cuda.synchronize()
    return out


# ------------------------------------------------------------------ golden vectors
def test_small_goldens_bit_exact(cuda, golden_small):
    g = golden_small
    for name in g["names"]:
        value, mask, want = g[f"{name}.value"], g[f"{name}.mask"], g[f"{name}.path"]
        got = run_mas(value, mask, cuda)             # default: per-cell mask, like the reference
        assert got.dtype == torch.from_numpy(want).dtype, name
        assert got.shape == want.shape
        assert np.array_equal(got.cpu().numpy(), want), name
        if name != "holes":                          # rectangular masks: the 8 B/cell path is exact too
            fast = run_mas(value, mask, cuda, strict_mask=False)
            assert fast.dtype == got.dtype and torch.equal(fast, got), name
        skew = run_mas(value, mask, cuda, flags=128)  # skewed-lane kernel (every dtype / staging mode)
        assert skew.dtype == got.dtype and torch.equal(skew, got), name
        if name != "holes":
            tma = run_mas(value, mask, cuda, strict_mask=False, flags=1 << 16)   # TMA tensor-load staging (fp32, aligned rows)
            assert tma.dtype == got.dtype and torch.equal(tma, got), name
        one = run_mas(value, mask, cuda, flags=1 << 18)   # one DP warp per utterance (what batches beyond one wave run)
        assert one.dtype == got.dtype and torch.equal(one, got), name


@pytest.mark.parametrize("flags", [0, 1 << 18, 1 << 16, 128, 1], ids=["lockstep", "lockstep_one_dp_warp", "lockstep_tma", "skewed", "general"])
def test_seeded_goldens_bit_exact(cuda, golden_seeded, flags):
    g = golden_seeded
    for name in g["names"]:
        seed, B, T_x, T_y = (int(v) for v in g[f"{name}.recipe"])
        if flags == 1 and B * T_x * T_y > 5_000_000:
            continue  # the general kernel is the slow fallback: keep its cases small
        value, t_x, t_y = seeded_case(seed, B, T_x, T_y, str(g[f"{name}.kind"]))
        path, dur, score = run_lengths(value, t_x, t_y, cuda, return_durations=True,
                                       return_score=True, flags=flags)
        assert np.array_equal(dur.cpu().numpy(), g[f"{name}.durations"]), name
        assert np.array_equal(score.cpu().numpy(), g[f"{name}.score"]), name
        p8 = path.to(torch.uint8).cpu().numpy()
        assert hashlib.sha256(p8.tobytes()).hexdigest() == str(g[f"{name}.path_sha256"]), name


# ------------------------------------------------------------------ oracle on random inputs
@pytest.mark.parametrize("flags", [0, 1 << 18, 1 << 16, 128, 1], ids=["lockstep", "lockstep_one_dp_warp", "lockstep_tma", "skewed", "general"])
def test_random_ragged_vs_oracle(cuda, flags):
    rng = np.random.default_rng(2024 + flags)
    for it in range(30):
        B = int(rng.integers(1, 7))
        T_x = int(rng.integers(1, 200))
        T_y = int(rng.integers(1, 420))
        t_x = rng.integers(0, T_x + 1, B).astype(np.int32)
        t_y = rng.integers(0, T_y + 1, B).astype(np.int32)     # degenerate + empty included
        value = (rng.integers(-3, 2, (B, T_x, T_y)) if it % 3 == 0 else
                 rng.standard_normal((B, T_x, T_y)) * 7).astype(np.float32)
        mask = rect_mask(t_x, t_y, T_x, T_y)
        tx_eff = mask.sum(1)[:, 0].astype(np.int32)   # what the reference derives from the mask
        ty_eff = mask.sum(2)[:, 0].astype(np.int32)
        want = oracle.maximum_path(value, mask)
        got = run_mas(value, mask, cuda)
        assert np.array_equal(got.cpu().numpy(), want), (it, B, T_x, T_y)
        path, dur = run_lengths(value, tx_eff, ty_eff, cuda, return_durations=True, flags=flags)
        assert np.array_equal(path.cpu().numpy(), want), (it, B, T_x, T_y, flags)
        assert np.array_equal(dur.cpu().numpy(), want.sum(-1).astype(np.int32))


@pytest.mark.parametrize("flags", [0, 1 << 18], ids=["auto", "one_dp_warp"])
@pytest.mark.parametrize("T_x", [1, 31, 32, 33, 64, 65, 190, 256, 257, 511, 512])
def test_token_axis_edges_vs_oracle(cuda, T_x, flags):
    rng = np.random.default_rng(T_x)
    T_y = T_x + int(rng.integers(0, 300))
    B = 3
    t_x = np.array([T_x, max(1, T_x - 1), max(1, T_x // 2)], np.int32)
    t_y = np.array([T_y, max(t_x[1], T_y - 17), t_x[2]], np.int32)   # last one: t_x == t_y
    value = (rng.standard_normal((B, T_x, T_y)) * 3 - 20).astype(np.float32)
    mask = rect_mask(t_x, t_y, T_x, T_y)
    want, wsc = oracle.maximum_path(value, mask, return_scores=True)
    path, dur, score = run_lengths(value, t_x, t_y, cuda, return_durations=True, return_score=True, flags=flags)
    assert np.array_equal(path.cpu().numpy(), want)
    assert np.array_equal(score.cpu().numpy(), wsc)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float64])
def test_value_dtypes_follow_reference_casts(cuda, dtype):
    rng = np.random.default_rng(11)
    B, T_x, T_y = 3, 45, 150
    t_x = np.array([45, 20, 33], np.int32)
    t_y = np.array([150, 100, 33], np.int32)
    v = torch.from_numpy(rng.standard_normal((B, T_x, T_y)) * 4).to(dtype)
    mask = torch.from_numpy(rect_mask(t_x, t_y, T_x, T_y)).to(dtype)
    # reference: (value*mask) in `dtype`, then astype(float32); path returned in `dtype`
    v32 = (v * mask).to(torch.float32).numpy()
    want = oracle.maximum_path(v32, mask.to(torch.float32).numpy())
    from art_tts_b200 import monotonic_align
    got = monotonic_align.maximum_path(v.to(cuda), mask.to(cuda))
    assert got.dtype == dtype
    assert np.array_equal(got.to(torch.float32).cpu().numpy(), want)


def test_mask_dtype_promotion_and_views(cuda):
    """mask may be bool / a broadcast view (x_mask[...,None] * y_mask[:,:,None] unmaterialised)."""
    from art_tts_b200 import monotonic_align
    rng = np.random.default_rng(3)
    B, T_x, T_y = 4, 50, 170
    t_x = torch.tensor([50, 10, 25, 49])
    t_y = torch.tensor([170, 60, 25, 169])
    value = torch.from_numpy(rng.standard_normal((B, T_x, T_y)).astype(np.float32)).to(cuda)
    xm = (torch.arange(T_x)[None, :] < t_x[:, None]).to(cuda)
    ym = (torch.arange(T_y)[None, :] < t_y[:, None]).to(cuda)
    view = xm[:, :, None] & ym[:, None, :]                      # bool
    want = oracle.maximum_path(value.cpu().numpy(), view.cpu().numpy().astype(np.float32))
    got = monotonic_align.maximum_path(value, view)
    assert got.dtype == torch.float32
    assert np.array_equal(got.cpu().numpy(), want)
    exp = xm[:, :, None].float().expand(B, T_x, T_y) * ym[:, None, :].float()
    got2 = monotonic_align.maximum_path(value, exp.transpose(1, 2).contiguous().transpose(1, 2))
    assert np.array_equal(got2.cpu().numpy(), want)


def test_out_dtypes_and_no_input_clobber(cuda):
    value, t_x, t_y = seeded_case(21, 4, 70, 260, "ljs")
    mask = rect_mask(t_x, t_y, 70, 260)
    want = oracle.maximum_path(value, mask)
    from art_tts_b200 import monotonic_align
    v = torch.from_numpy(value).to(cuda)
    keep = v.clone()
    for dt in (torch.float32, torch.float16, torch.bfloat16, torch.float64, torch.int32, torch.uint8):
        got = monotonic_align.maximum_path_lengths(v, torch.from_numpy(t_x), torch.from_numpy(t_y),
                                                   out_dtype=dt)
        assert got.dtype == dt
        assert np.array_equal(got.to(torch.float32).cpu().numpy(), want), dt
    assert torch.equal(v, keep)   # unlike the reference, the input is never modified


# ------------------------------------------------------------------ full-size properties
def check_path_invariants(path, dur, t_x, t_y):
    """SURVEY.md App. A.4 for 1 <= t_x <= t_y, on device."""
    B, T_x, T_y = path.shape
    tx = torch.as_tensor(t_x, device=path.device).long()
    ty = torch.as_tensor(t_y, device=path.device).long()
    col = path.sum(1)                                   # ones per frame
    frames = torch.arange(T_y, device=path.device)[None, :]
    assert torch.equal(col, (frames < ty[:, None]).to(col.dtype))
    assert torch.equal(path.sum(2).to(torch.int32), dur)
    tokens = torch.arange(T_x, device=path.device)[None, :]
    valid = tokens < tx[:, None]
    assert bool(((dur >= 1) == valid).all())            # every valid token >= 1 frame, padding 0
    assert torch.equal(dur.sum(1).long(), ty)
    # monotone: token index per frame is non-decreasing and moves by at most 1
    idx = path.argmax(1)
    step = idx[:, 1:] - idx[:, :-1]
    inside = frames[:, 1:] < ty[:, None]
    assert bool(((step == 0) | (step == 1))[inside].all())


def test_config1_ljspeech_batch_bit_exact_and_invariants(cuda):
    value, t_x, t_y = seeded_case(0, 16, 190, 870, "ljs")
    mask = rect_mask(t_x, t_y, 190, 870)
    want = oracle.maximum_path(value, mask, n_threads=8)
    path, dur = run_lengths(value, t_x, t_y, cuda, return_durations=True)
    assert np.array_equal(path.cpu().numpy(), want)
    check_path_invariants(path, dur, t_x, t_y)


def test_config4_long_utterances_spilled_bits(cuda):
    """T_x=512, T_y=4096: direction bits (262 KB/utterance) spill to the workspace."""
    value, t_x, t_y = seeded_case(4, 4, 512, 4096, "full")
    want = oracle.maximum_path(value, rect_mask(t_x, t_y, 512, 4096), n_threads=4)
    path, dur = run_lengths(value, t_x, t_y, cuda, return_durations=True)
    assert np.array_equal(path.cpu().numpy(), want)
    check_path_invariants(path, dur, t_x, t_y)


def test_config5_b1024_properties_and_generate_path_roundtrip(cuda):
    """BASELINE config 5 size: checked through size-independent properties + a sampled oracle."""
    from art_tts_b200 import monotonic_align, utils
    B, T_x, T_y = 1024, 190, 870
    g = torch.Generator(device="cpu").manual_seed(5)
    t_x = torch.randint(60, T_x + 1, (B,), generator=g, dtype=torch.int32)
    t_y = torch.minimum(torch.tensor(T_y), 4 * t_x + torch.randint(0, 100, (B,), generator=g,
                                                                    dtype=torch.int32)).int()
    torch.manual_seed(5)
    value = -(torch.rand(B, T_x, T_y, device=cuda) * 100 + 50)
    path, dur, score = monotonic_align.maximum_path_lengths(value, t_x, t_y, return_durations=True,
                                                            return_score=True)
    check_path_invariants(path, dur, t_x.numpy(), t_y.numpy())
    # score == sum of value over the path (linearity of the objective along the chosen path)
    s2 = (value.double() * path.double()).sum((1, 2))
    assert torch.allclose(score.double(), s2, rtol=1e-5)
    # generate_path is the exact inverse (utils.py:26-43)
    again = utils.generate_path_lengths(dur, t_x, t_y, T_y, out_dtype=torch.float32)
    assert torch.equal(again, path)
    # idempotent / deterministic
    path2 = monotonic_align.maximum_path_lengths(value, t_x, t_y)
    assert torch.equal(path2, path)
    # sampled utterances against the oracle, bit-exact
    pick = [0, 1, 17, 500, 1023]
    sub = value[pick].cpu().numpy()
    want = oracle.maximum_path(sub, rect_mask(t_x[pick].numpy(), t_y[pick].numpy(), T_x, T_y))
    assert np.array_equal(path[pick].cpu().numpy(), want)


# ------------------------------------------------------------------ lengths / generate_path
def test_lengths_from_mask_dtypes(cuda):
    from art_tts_b200 import monotonic_align
    t_x = np.array([7, 1, 0, 5], np.int32)
    t_y = np.array([13, 2, 0, 9], np.int32)
    m = rect_mask(t_x, t_y, 9, 13)
    for dt in (torch.float32, torch.float16, torch.bfloat16, torch.float64, torch.int32,
               torch.int64, torch.uint8, torch.bool):
        lx, ly = monotonic_align.lengths_from_mask(torch.from_numpy(m).to(dt).to(cuda))
        assert lx.cpu().tolist() == (m.sum(1)[:, 0]).astype(int).tolist(), dt
        assert ly.cpu().tolist() == (m.sum(2)[:, 0]).astype(int).tolist(), dt


def test_generate_path_matches_reference_formula(cuda):
    from art_tts_b200 import utils
    rng = np.random.default_rng(9)
    B, T_x, T_y = 5, 37, 211
    t_x = np.array([37, 20, 1, 36, 9], np.int32)
    t_y = np.array([211, 90, 5, 210, 64], np.int32)
    mask = rect_mask(t_x, t_y, T_x, T_y)
    dur = rng.integers(0, 9, (B, T_x)).astype(np.float32)         # fp32 like ceil(w) (tts.py:132)
    want = oracle.generate_path(dur, mask)
    got = utils.generate_path(torch.from_numpy(dur).to(cuda), torch.from_numpy(mask).to(cuda))
    assert got.dtype == torch.float32
    assert np.array_equal(got.cpu().numpy(), want)
    frac = (dur * 1.3).astype(np.float32)                          # length_scale != 1
    want = oracle.generate_path(frac, mask)
    got = utils.generate_path(torch.from_numpy(frac).to(cuda), torch.from_numpy(mask).to(cuda))
    assert np.array_equal(got.cpu().numpy(), want)
    got = utils.generate_path(torch.from_numpy(dur.astype(np.int32)).to(cuda),
                              torch.from_numpy(mask).to(cuda))
    assert np.array_equal(got.cpu().numpy(), oracle.generate_path(dur.astype(np.int32), mask))


def test_generate_path_matches_reference_goldens(cuda):
    """generate_path_kernel against the reference's own model.utils.generate_path (utils.py:26-43)
    on its inference call site's inputs (tts.py:130-147): fp32 durations with length_scale
    1.0 / 1.3 / 0.77 and the integer durations MAS produces (tests/golden/generate_path.npz)."""
    from art_tts_b200 import utils
    g = np.load(os.path.join(GOLDEN, "generate_path.npz"))
    for name in g["names"]:
        dur, T_y = g[f"{name}.duration"], int(g[f"{name}.T_y"])
        mask = rect_mask(g[f"{name}.x_lengths"], g[f"{name}.y_lengths"], dur.shape[1], T_y)
        want = np.unpackbits(g[f"{name}.path_packed"], axis=-1)[:, :, :T_y].astype(np.float32)
        got = utils.generate_path(torch.from_numpy(dur).to(cuda), torch.from_numpy(mask).to(cuda))
        assert got.dtype == torch.float32, name
        assert np.array_equal(got.cpu().numpy(), want), name


def test_config3_size_reference_capture(cuda):
    """BASELINE config 3 at its stated size: mu_x captured from the reference's GradTTS(params_v2)
    encoder, B=64, T_x<=190, T_y<=872 (tests/golden/cfg3_gradtts.npz).  Fused kernel (both engines)
    against the reference's log-prior rows, alignment and losses; drop-in on the kernel's own prior
    bit-exact against the oracle."""
    from art_tts_b200 import alignment, monotonic_align
    g = np.load(os.path.join(GOLDEN, "cfg3_gradtts.npz"))
    _, x_len, y, y_len = cfg3_inputs()
    assert hashlib.sha256(y.tobytes()).hexdigest() == str(g["y_sha256"])
    mu_x = g["mu_x"]
    B, F, T_x = mu_x.shape
    T_y = y.shape[2]
    mask = rect_mask(x_len, y_len, T_x, T_y)
    want_path = path_from_durations(g["durations"], y_len, T_y)
    assert hashlib.sha256(want_path.tobytes()).hexdigest() == str(g["path_sha256"])
    rows = g["log_prior_rows"]
    m = mask[:, ::19, :].astype(bool)
    for engine in ("tensor_core", "cuda_core"):
        path, dur, score, lp = fused(mu_x, y, x_len, y_len, cuda, return_score=True,
                                     return_log_prior=True, flags=ENGINES[engine])
        lp_np = lp.cpu().numpy()
        rel = np.abs(lp_np[:, ::19, :] - rows)[m] / np.abs(rows[m])
        assert rel.max() <= 1e-5, (engine, rel.max())                       # prior within 1e-5 relative
        agree = (path.to(torch.uint8).cpu().numpy() == want_path).mean()
        assert agree >= 0.999, (engine, agree)                              # >= 99.9 % of cells
        assert np.allclose(score.cpu().numpy(), g["score"], rtol=1e-4), engine   # log-likelihood 1e-4 rel
        tok = (dur.cpu().numpy() == g["durations"]).mean()
        assert tok >= 0.99, (engine, tok)
        # drop-in on the kernel's own prior: bit-exact against the oracle on the same tensor
        got = monotonic_align.maximum_path(lp, torch.from_numpy(mask).to(cuda))
        assert np.array_equal(got.cpu().numpy(), oracle.maximum_path(lp_np, mask, n_threads=8)), engine
        assert torch.equal(got, path), engine
    # duration loss on the reference's durations (tts.py:503-506)
    d = torch.from_numpy(g["durations"]).to(cuda)
    logw = torch.from_numpy(g["logw"]).to(cuda)
    loss = alignment.duration_loss_from_durations(logw, d, torch.from_numpy(x_len.astype(np.int32)))
    assert np.isclose(float(loss), float(g["dur_loss"]), rtol=1e-5)


@pytest.mark.parametrize("cl", [0, 1 << 17], ids=["cluster4", "cluster2"])
def test_long_text_reference_capture_on_clusters(cuda, cl):
    """Token axis of 257..420 (add-blank text; what BASELINE config 4 stands for): mu_x captured from the
    reference's GradTTS(params_v2) encoder, alignment and losses from its own compute_loss
    (tests/golden/long_text_gradtts.npz).  The fused kernel runs this shape on thread-block clusters."""
    from art_tts_b200 import _lib, alignment
    g = np.load(os.path.join(GOLDEN, "long_text_gradtts.npz"))
    _, x_len, y, y_len = long_text_inputs()
    assert hashlib.sha256(y.tobytes()).hexdigest() == str(g["y_sha256"])
    mu_x = g["mu_x"]
    B, F, T_x = mu_x.shape
    T_y = y.shape[2]
    assert _lib.load().mas_from_prior_plan(B, F, T_x, T_y, cl) == 0
    mask = rect_mask(x_len, y_len, T_x, T_y)
    want_path = path_from_durations(g["durations"], y_len, T_y)
    assert hashlib.sha256(want_path.tobytes()).hexdigest() == str(g["path_sha256"])
    st = int(g["row_step"])
    rows = g["log_prior_rows"]
    m = mask[:, ::st, :].astype(bool)
    path, dur, score, lp = fused(mu_x, y, x_len, y_len, cuda, return_score=True, return_log_prior=True, flags=cl)
    lp_np = lp.cpu().numpy()
    rel = np.abs(lp_np[:, ::st, :] - rows)[m] / np.abs(rows[m])
    assert rel.max() <= 1e-5, rel.max()                                     # prior within 1e-5 relative
    agree = (path.to(torch.uint8).cpu().numpy() == want_path).mean()
    assert agree >= 0.999, agree                                            # >= 99.9 % of cells
    assert np.allclose(score.cpu().numpy(), g["score"], rtol=1e-4)          # log-likelihood 1e-4 rel
    assert (dur.cpu().numpy() == g["durations"]).mean() >= 0.99
    self_path = oracle.maximum_path(lp_np, mask, n_threads=8)
    assert np.array_equal(path.cpu().numpy(), self_path)                    # bit-exact on the kernel's own prior
    d = torch.from_numpy(g["durations"]).to(cuda)
    loss = alignment.duration_loss_from_durations(torch.from_numpy(g["logw"]).to(cuda), d,
                                                  torch.from_numpy(x_len.astype(np.int32)))
    assert np.isclose(float(loss), float(g["dur_loss"]), rtol=1e-5)


# ------------------------------------------------------------------ fused prior + MAS
def fused(mu_x, y, x_len, y_len, cuda, **kw):
    from art_tts_b200 import monotonic_align
    out = monotonic_align.maximum_path_from_prior(
        torch.from_numpy(mu_x).to(cuda), None, torch.from_numpy(y).to(cuda),
        torch.from_numpy(np.asarray(x_len, np.int32)), torch.from_numpy(np.asarray(y_len, np.int32)),
        **kw)
    torch.cuda.synchronize()
    return out


def prior_bars(lp_gpu, path_gpu, score_gpu, lp_ref, path_ref, mask):
    m = mask.astype(bool)
    rel = np.abs(lp_gpu - lp_ref)[m] / np.maximum(np.abs(lp_ref[m]), 1e-6)
    assert rel.max() <= 1e-5, rel.max()                       # prior within 1e-5 relative
    agree = (path_gpu == path_ref).mean()
    assert agree >= 0.999, agree                              # >= 99.9 % of cells
    ll_ref = (lp_ref.astype(np.float64) * path_ref).sum((1, 2))
    assert np.allclose(score_gpu, ll_ref, rtol=1e-4)          # total log-likelihood 1e-4 relative


@pytest.mark.parametrize("fname", ["prior_gradtts.npz", "prior_arttts.npz"])
def test_fused_matches_reference_model_capture(cuda, fname):
    """mu_x / y captured inside the reference's compute_loss (tts.py:472-499)."""
    g = np.load(os.path.join(GOLDEN, fname))
    mu_x, y = g["mu_x"], g["y"]
    x_len, y_len = g["x_lengths"], g["y_lengths"]
    T_x, T_y = mu_x.shape[2], y.shape[2]
    mask = rect_mask(x_len, y_len, T_x, T_y)
    path, dur, score, fidx, lp = fused(mu_x, y, x_len, y_len, cuda, return_score=True,
                                       return_frame_idx=True, return_log_prior=True)
    want_path = np.unpackbits(g["attn_packed"], axis=-1)[:, :, :T_y].astype(np.float32)
    prior_bars(lp.cpu().numpy(), path.cpu().numpy(), score.cpu().numpy(), g["log_prior"], want_path,
               mask)
    # the kernel's own prior -> oracle MAS must reproduce the kernel's path bit-exactly
    self_path = oracle.maximum_path(lp.cpu().numpy(), mask)
    assert np.array_equal(path.cpu().numpy(), self_path)
    assert np.array_equal(dur.cpu().numpy(), self_path.sum(-1).astype(np.int32))
    fi = fidx.cpu().numpy()
    for b in range(mu_x.shape[0]):
        assert np.array_equal(fi[b, :y_len[b]], self_path[b].argmax(0)[:y_len[b]])
        assert (fi[b, y_len[b]:] == -1).all()


# MAS_FLAG_NO_TENSOR / MAS_FLAG_FORCE_TENSOR (/ + MAS_FLAG_STAGGER_MMA: the other MMA issue order of the tensor-core engine)
ENGINES = {"auto": 0, "cuda_core": 16, "tensor_core": 32, "tensor_core_staggered": 32 | (1 << 20)}


def select_engine(monkeypatch, engine):
    return ENGINES[engine]


@pytest.mark.parametrize("engine", list(ENGINES))
@pytest.mark.parametrize("F,B,T_x,T_y,seed", [(16, 32, 160, 512, 1), (80, 16, 190, 870, 2),
                                              (80, 3, 33, 95, 3), (7, 2, 5, 40, 4)])
def test_fused_vs_oracle_configs(cuda, F, B, T_x, T_y, seed, engine, monkeypatch):
    """BASELINE config 2 (articulatory, F=16, ragged) and the LJSpeech shape (F=80), through the
    engines of the fused kernel: fp32 FMA on CUDA cores and 3xTF32 on the tensor cores."""
    select_engine(monkeypatch, engine)
    rng = np.random.default_rng(seed)
    x_len = rng.integers(max(1, T_x // 8), T_x + 1, B).astype(np.int32)
    y_len = np.minimum(T_y, 3 * x_len + rng.integers(0, 61, B)).astype(np.int32)
    x_len[0], y_len[0] = T_x, T_y
    mu_x = rng.standard_normal((B, F, T_x)).astype(np.float32)
    y = rng.standard_normal((B, F, T_y)).astype(np.float32)
    if F == 16:
        y[:, [12, 14]] = 0.0
    mu_x *= (np.arange(T_x)[None, None, :] < x_len[:, None, None])
    y *= (np.arange(T_y)[None, None, :] < y_len[:, None, None])
    mask = rect_mask(x_len, y_len, T_x, T_y)
    lp_ref = oracle.log_prior(mu_x, y)
    path_ref = oracle.maximum_path(lp_ref, mask, n_threads=8)
    path, dur, score, lp = fused(mu_x, y, x_len, y_len, cuda, return_score=True,
                                 return_log_prior=True, flags=ENGINES[engine])
    prior_bars(lp.cpu().numpy(), path.cpu().numpy(), score.cpu().numpy(), lp_ref, path_ref, mask)
    self_path = oracle.maximum_path(lp.cpu().numpy(), mask, n_threads=8)
    assert np.array_equal(path.cpu().numpy(), self_path)
    assert np.array_equal(dur.cpu().numpy(), self_path.sum(-1).astype(np.int32))
    # without the tap (band-only production of the prior) the path must not change
    p2, d2 = fused(mu_x, y, x_len, y_len, cuda, flags=ENGINES[engine])
    assert torch.equal(p2, path) and torch.equal(d2, dur)


@pytest.mark.parametrize("engine", ["tensor_core"])
@pytest.mark.parametrize("T_x", [1, 31, 32, 33, 63, 64, 65, 127, 128, 129, 191, 192, 255, 256, 257])
def test_tensor_core_engine_token_axis_edges(cuda, T_x, engine, monkeypatch):
    """Token counts around every ownership boundary of the tensor-core kernel: 32 lanes x 2 DP
    warps (64), the 128-token M tile, and its 256-token limit (257 falls back to CUDA cores)."""
    select_engine(monkeypatch, engine)
    rng = np.random.default_rng(100 + T_x)
    B, F = 5, 24
    T_y = max(T_x + 40, 3 * T_x // 2)
    x_len = np.array([T_x, max(1, T_x - 1), max(1, T_x // 2), 1, max(1, T_x - 33)], np.int32)
    y_len = np.array([T_y, T_y - 7, max(1, T_x // 2), 33, T_y - 1], np.int32)   # incl. t_x == t_y
    mu_x = rng.standard_normal((B, F, T_x)).astype(np.float32)
    y = rng.standard_normal((B, F, T_y)).astype(np.float32)
    mask = rect_mask(x_len, y_len, T_x, T_y)
    path, dur, score, fidx, lp = fused(mu_x, y, x_len, y_len, cuda, return_score=True,
                                       return_frame_idx=True, return_log_prior=True,
                                       flags=ENGINES["tensor_core"])
    lp_ref = oracle.log_prior(mu_x, y)
    prior_bars(lp.cpu().numpy(), path.cpu().numpy(), score.cpu().numpy(), lp_ref,
               oracle.maximum_path(lp_ref, mask), mask)
    self_path = oracle.maximum_path(lp.cpu().numpy(), mask)
    assert np.array_equal(path.cpu().numpy(), self_path)
    assert np.array_equal(dur.cpu().numpy(), self_path.sum(-1).astype(np.int32))
    assert np.array_equal(fidx.cpu().numpy(), oracle.frame_index(self_path, y_len))


def test_fused_engines_agree_at_b1024(cuda):
    """Bench shape: both engines see the same inputs; their priors differ by rounding only, so
    durations agree on >= 99.9 % of tokens and every invariant of the path holds for both."""
    B, F, T_x, T_y = 1024, 80, 190, 872
    rng = np.random.default_rng(7)
    x_len = rng.integers(60, T_x + 1, B).astype(np.int32)
    y_len = np.minimum(870, 4 * x_len + rng.integers(0, 100, B)).astype(np.int32)
    gen = torch.Generator(device=cuda).manual_seed(3)
    mu_x = torch.randn(B, F, T_x, device=cuda, generator=gen)
    y = torch.randn(B, F, T_y, device=cuda, generator=gen)
    from art_tts_b200 import monotonic_align
    tx, ty = torch.from_numpy(x_len).to(cuda), torch.from_numpy(y_len).to(cuda)
    outs = {}
    for name, fl in ENGINES.items():
        path, dur, score = monotonic_align.maximum_path_from_prior(mu_x, None, y, tx, ty, return_score=True,
                                                                   flags=fl)
        assert torch.equal(path.sum(1)[torch.arange(T_y, device=cuda)[None] < ty[:, None]],
                           torch.ones(int(ty.sum()), device=cuda))            # one token per valid frame
        assert torch.equal(dur.sum(1), ty) and torch.equal(path.sum(2).int(), dur)
        outs[name] = (dur, score)
    agree = (outs["cuda_core"][0] == outs["tensor_core"][0]).float().mean().item()
    assert agree >= 0.999, agree
    assert torch.allclose(outs["cuda_core"][1], outs["tensor_core"][1], rtol=1e-5)
    assert torch.equal(outs["auto"][0], outs["tensor_core"][0])                  # F = 80 -> tensor cores


def test_fused_unfused_plan_for_long_text(cuda):
    """T_x=512 at F=80 on the CUDA-core engine: mu_x does not fit beside the ring -> prior to HBM once, then
    drop-in (the tensor-core engine runs this shape on a cluster of CTAs, see below)."""
    from art_tts_b200 import _lib
    rng = np.random.default_rng(6)
    B, F, T_x, T_y = 2, 80, 512, 1200
    assert _lib.load().mas_from_prior_plan(B, F, T_x, T_y, ENGINES["cuda_core"]) == 1
    assert _lib.load().mas_from_prior_plan(B, F, T_x, T_y, 0) == 0
    mu_x = rng.standard_normal((B, F, T_x)).astype(np.float32)
    y = rng.standard_normal((B, F, T_y)).astype(np.float32)
    x_len = np.array([512, 300], np.int32)
    y_len = np.array([1200, 1000], np.int32)
    mask = rect_mask(x_len, y_len, T_x, T_y)
    path, dur, lp = fused(mu_x, y, x_len, y_len, cuda, return_log_prior=True, flags=ENGINES["cuda_core"])
    self_path = oracle.maximum_path(lp.cpu().numpy(), mask, n_threads=2)
    assert np.array_equal(path.cpu().numpy(), self_path)


CLUSTERS = {"cluster4": 0, "cluster2": 1 << 17}


@pytest.mark.parametrize("cl", list(CLUSTERS))
@pytest.mark.parametrize("T_x", [257, 300, 384, 385, 449, 512])
def test_fused_cluster_token_axis_edges(cuda, T_x, cl):
    """256 < T_x <= 512: one thread-block cluster per utterance, the token axis split over its CTAs (4 x 128
    or 2 x 256 tokens), the recurrence and the backtrack crossing CTAs through distributed shared memory.
    Utterances that end in every CTA of the cluster, t_x == t_y, empty and degenerate ones in one batch."""
    rng = np.random.default_rng(500 + T_x)
    F = 24
    T_y = 3 * T_x // 2 + 37
    x_len = np.array([T_x, T_x - 1, 128, 129, 256, min(T_x, 257), 1, 200, 0, T_x, 60, max(1, T_x - 130)], np.int32)
    y_len = np.array([T_y, T_y - 7, 128, T_y, 300, T_y - 1, 33, T_y, 50, T_x - 5, 61, T_y - 64], np.int32)
    B = len(x_len)
    mu_x = rng.standard_normal((B, F, T_x)).astype(np.float32)
    y = rng.standard_normal((B, F, T_y)).astype(np.float32)
    mask = rect_mask(x_len, y_len, T_x, T_y)
    from art_tts_b200 import _lib
    assert _lib.load().mas_from_prior_plan(B, F, T_x, T_y, ENGINES["tensor_core"] | CLUSTERS[cl]) == 0
    path, dur, score, fidx, lp = fused(mu_x, y, x_len, y_len, cuda, return_score=True,
                                       return_frame_idx=True, return_log_prior=True,
                                       flags=ENGINES["tensor_core"] | CLUSTERS[cl])
    lp_np = lp.cpu().numpy()
    ref_lp = oracle.log_prior(mu_x, y, method="c")
    m = mask.astype(bool)
    assert np.allclose(lp_np[m], ref_lp[m], rtol=1e-5, atol=1e-4)
    # the degenerate utterance (t_x > t_y) is decided by raw values computed with the CUDA-core formula
    lp_for_oracle = lp_np.copy()
    lp_for_oracle[9] = ref_lp[9]
    want = oracle.maximum_path(lp_for_oracle, mask, n_threads=8)
    assert np.array_equal(path.cpu().numpy(), want)
    assert np.array_equal(dur.cpu().numpy(), want.sum(-1).astype(np.int32))
    fi = fidx.cpu().numpy()
    for b in range(B):
        if 1 <= x_len[b] <= y_len[b]:
            assert np.array_equal(fi[b, :y_len[b]], want[b].argmax(0)[:y_len[b]]), b
            assert (fi[b, y_len[b]:] == -1).all(), b
    ok = (x_len >= 1) & (x_len <= y_len)
    ll = (lp_np.astype(np.float64) * want).sum((1, 2))
    assert np.allclose(score.cpu().numpy()[ok], ll[ok], rtol=1e-4)
    # without the tap (band-only production of the prior) nothing changes; nor with several utterances per cluster
    for fl in (0, 3 << 8):
        p2, d2 = fused(mu_x, y, x_len, y_len, cuda, flags=ENGINES["tensor_core"] | CLUSTERS[cl] | fl)
        assert torch.equal(p2, path) and torch.equal(d2, dur), fl


@pytest.mark.timeout(300)
@pytest.mark.parametrize("cl", list(CLUSTERS))
def test_fused_cluster_random_batches(cuda, cl):
    """Random ragged batches on the cluster kernel: utterances of every length class in random order (so the
    links between neighbouring CTAs switch on and off from one utterance to the next), several utterances per
    cluster, small and large frame counts; equal to the oracle on the kernel's own prior every time."""
    rng = np.random.default_rng(77 if cl == "cluster4" else 78)
    for it in range(12):
        T_x = int(rng.integers(257, 513))
        T_y = int(rng.integers(T_x, 3 * T_x))
        B = int(rng.integers(1, 48))
        F = int(rng.choice([32, 40, 80]))
        x_len = rng.integers(1, T_x + 1, B).astype(np.int32)
        x_len[rng.integers(0, B)] = T_x
        y_len = np.minimum(T_y, x_len + rng.integers(0, 2 * T_x, B)).astype(np.int32)
        mu_x = rng.standard_normal((B, F, T_x)).astype(np.float32)
        y = rng.standard_normal((B, F, T_y)).astype(np.float32)
        mask = rect_mask(x_len, y_len, T_x, T_y)
        upc = int(rng.choice([0, 2, 5]))
        st = (1 << 20) if it % 2 else 0          # every other batch: the staggered MMA issue order
        path, dur, score, lp = fused(mu_x, y, x_len, y_len, cuda, return_score=True, return_log_prior=True,
                                     flags=CLUSTERS[cl] | (upc << 8) | st)
        lp_np = lp.cpu().numpy()
        want, wsc = oracle.maximum_path(lp_np, mask, n_threads=8, return_scores=True)
        assert np.array_equal(path.cpu().numpy(), want), (it, T_x, T_y, B, F, upc)
        assert np.array_equal(dur.cpu().numpy(), want.sum(-1).astype(np.int32)), (it, T_x, T_y, B, F, upc)
        assert np.allclose(score.cpu().numpy(), wsc, rtol=1e-5), (it, T_x, T_y, B, F, upc)
        p2, d2 = fused(mu_x, y, x_len, y_len, cuda, flags=CLUSTERS[cl] | (upc << 8) | st)
        assert torch.equal(p2, path) and torch.equal(d2, dur), (it, T_x, T_y, B, F, upc)


@pytest.mark.parametrize("cl", list(CLUSTERS))
def test_config4_fused_on_clusters(cuda, cl):
    """BASELINE config 4 (T_text=512, T_mel=4096, F=80) through the fused entry: the prior is never written to
    HBM (plan 0), direction words go through the workspace, result equal to the oracle on the kernel's own prior."""
    from art_tts_b200 import _lib
    rng = np.random.default_rng(44)
    B, F, T_x, T_y = 6, 80, 512, 4096
    assert _lib.load().mas_from_prior_plan(32, F, T_x, T_y, 0) == 0
    x_len = np.array([512, 512, 400, 511, 130, 300], np.int32)
    y_len = np.array([4096, 3000, 4096, 2047, 4001, 1111], np.int32)
    mu_x = rng.standard_normal((B, F, T_x)).astype(np.float32)
    y = rng.standard_normal((B, F, T_y)).astype(np.float32)
    mask = rect_mask(x_len, y_len, T_x, T_y)
    path, dur, lp = fused(mu_x, y, x_len, y_len, cuda, return_log_prior=True, flags=CLUSTERS[cl])
    lp_np = lp.cpu().numpy()
    ref_lp = oracle.log_prior(mu_x, y, method="c")
    m = mask.astype(bool)
    assert np.allclose(lp_np[m], ref_lp[m], rtol=1e-5, atol=1e-4)
    want = oracle.maximum_path(lp_np, mask, n_threads=8)
    assert np.array_equal(path.cpu().numpy(), want)
    assert np.array_equal(dur.cpu().numpy(), want.sum(-1).astype(np.int32))
    p2, d2 = fused(mu_x, y, x_len, y_len, cuda, flags=CLUSTERS[cl])
    assert torch.equal(p2, path) and torch.equal(d2, dur)
    check_path_invariants(path, dur, x_len, y_len)


@pytest.mark.parametrize("chunk,trim", [(0, True), (5, True), (64, False)])
def test_host_buffer_entry_equals_device_entry(cuda, chunk, trim):
    """mas_from_prior_host_f32 (trimmed, chunked H2D overlapped with the kernels) must give
    exactly what mas_from_prior_f32 gives on the same batch already resident in HBM."""
    from art_tts_b200 import _lib, monotonic_align
    B, F, T_x, T_y = 37, 80, 70, 300
    rng = np.random.default_rng(21)
    x_len = rng.integers(5, T_x + 1, B).astype(np.int32)
    y_len = np.minimum(T_y, 4 * x_len + rng.integers(0, 20, B)).astype(np.int32)
    order = np.argsort(-(x_len.astype(np.int64) * y_len), kind="stable")
    x_len, y_len = x_len[order], y_len[order]
    mu_x = rng.standard_normal((B, F, T_x)).astype(np.float32)
    y = rng.standard_normal((B, F, T_y)).astype(np.float32)
    h = [torch.from_numpy(a).pin_memory() for a in (mu_x, y, x_len, y_len)]
    h_dur = torch.zeros(B, T_x, dtype=torch.int32).pin_memory()
    h_score = torch.zeros(B, dtype=torch.float32).pin_memory()
    # poison the staging buffers first so that stale padding would show up
    monotonic_align.maximum_path_from_prior_host(
        torch.full_like(h[0], float("nan")), torch.full_like(h[1], float("nan")), h[2], h[3], cuda)
    path, dur, score, moved = monotonic_align.maximum_path_from_prior_host(
        h[0], h[1], h[2], h[3], cuda, chunk=chunk, durations_host=h_dur, score_host=h_score,
        flags=0 if trim else _lib.FLAG_HOST_NO_TRIM)
    torch.cuda.synchronize()
    p2, d2, s2 = monotonic_align.maximum_path_from_prior(
        h[0].to(cuda), None, h[1].to(cuda), h[2], h[3], return_score=True)
    assert torch.equal(path, p2) and torch.equal(dur, d2) and torch.equal(score, s2)
    assert torch.equal(h_dur, d2.cpu()) and torch.equal(h_score, s2.cpu())
    full = (mu_x.nbytes + y.nbytes + 8 * B)
    assert moved == full if not trim else (moved < full if chunk else moved <= full)


def test_entry_points_are_cuda_graph_capturable_and_stream_ordered(cuda):
    """The C ABI promises: asynchronous on the caller's stream, no host sync, graph capturable.
    Capture drop-in + fused + the alignment consumers in one graph on a side stream and replay it
    on fresh inputs; results must equal the eager ones."""
    from art_tts_b200 import alignment, monotonic_align
    B, F, T_x, T_y = 8, 80, 50, 200
    rng = np.random.default_rng(9)
    x_len = torch.from_numpy(rng.integers(10, T_x + 1, B).astype(np.int32)).to(cuda)
    y_len = torch.from_numpy(rng.integers(120, T_y + 1, B).astype(np.int32)).to(cuda)
    mu_x = torch.randn(B, F, T_x, device=cuda)
    y = torch.randn(B, F, T_y, device=cuda)
    value = torch.randn(B, T_x, T_y, device=cuda)

    def run():
        p1, d1 = monotonic_align.maximum_path_lengths(value, x_len, y_len, return_durations=True)
        p2, d2, fi = monotonic_align.maximum_path_from_prior(mu_x, None, y, x_len, y_len, return_frame_idx=True)
        mu_y = alignment.aligned_mu_y(mu_x, fi, None, y_len)
        return p1, d1, p2, d2, mu_y

    run()                                    # warm-up: workspaces, function attributes
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            outs = run()
    # new inputs, same buffers
    mu_x.normal_()
    y.normal_()
    value.normal_()
    graph.replay()
    torch.cuda.synchronize()
    want = run()
    torch.cuda.synchronize()
    for a, b in zip(outs, want):
        assert torch.equal(a, b)


@pytest.mark.parametrize("engine", ["cuda_core", "tensor_core"])
def test_fused_degenerate_and_empty_utterances(cuda, engine, monkeypatch):
    """Empty utterances (t_x == 0 or t_y == 0: all-zero path, like the reference's all-zero mask),
    t_x == t_y (pure diagonal), t_x == 1, and the reference's degenerate t_x > t_y case (backtrack
    over raw prior values) in the middle of a batch of normal ones."""
    select_engine(monkeypatch, engine)
    rng = np.random.default_rng(31)
    B, F, T_x, T_y = 8, 40, 20, 64
    x_len = np.array([0, 5, 9, 1, 7, 3, 20, 12], np.int32)
    y_len = np.array([10, 0, 5, 1, 7, 50, 64, 12], np.int32)
    mu_x = rng.standard_normal((B, F, T_x)).astype(np.float32)
    y = rng.standard_normal((B, F, T_y)).astype(np.float32)
    mask = rect_mask(x_len, y_len, T_x, T_y)
    path, dur, score, lp = fused(mu_x, y, x_len, y_len, cuda, return_score=True, return_log_prior=True,
                                 flags=ENGINES[engine])
    lp_np, path_np = lp.cpu().numpy(), path.cpu().numpy()
    # the degenerate utterance (b = 2) is decided by raw values both kernels compute with the
    # CUDA-core formula; the tap of the tensor-core engine differs from it by rounding only, so
    # feed the oracle the fp32 FMA prior there
    ref_lp = oracle.log_prior(mu_x, y, method="c")
    lp_for_oracle = lp_np.copy()
    lp_for_oracle[2] = ref_lp[2]
    want = oracle.maximum_path(lp_for_oracle, mask)
    assert np.array_equal(path_np, want)
    assert np.array_equal(dur.cpu().numpy(), want.sum(-1).astype(np.int32))
    assert path_np[0].sum() == 0 and path_np[1].sum() == 0
    assert np.array_equal(path_np[4, :7, :7], np.eye(7, dtype=np.float32))
    m = mask.astype(bool)
    assert np.allclose(lp_np[m], ref_lp[m], rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("T_x,T_y", [(70, 300), (300, 700)], ids=["one_cta", "cluster"])
def test_peer_gather_single_rank_and_refusal(cuda, T_x, T_y):
    """mas_peer_gather with this GPU's own buffers as the only peer (the multi-GPU form is
    tests/test_gpu_multi.py): the tensor-core engine writes durations and frame index at row0 + b
    through the row strides, the CUDA-core engine refuses the call instead of leaving the buffers
    unwritten, a call that does not fit the buffers is refused, and without a description nothing
    is written (no process-wide state)."""
    import ctypes
    from art_tts_b200 import _lib
    rng = np.random.default_rng(5)
    B, F = 40, 80
    x_len = rng.integers(5, T_x + 1, B).astype(np.int32)
    y_len = np.minimum(T_y, (4 if T_x < 256 else 2) * x_len + rng.integers(0, 20, B)).astype(np.int32)
    x_len = np.minimum(x_len, y_len)
    mu_x = rng.standard_normal((B, F, T_x)).astype(np.float32)
    y = rng.standard_normal((B, F, T_y)).astype(np.float32)
    row0, stride, fstride = 3, T_x + 9, T_y + 4        # buffers padded to a larger T_x / T_y
    buf = torch.full((row0 + B + 2, stride), -7, dtype=torch.int32, device=cuda)
    fbuf = torch.full((row0 + B + 2, fstride), -7, dtype=torch.int32, device=cuda)
    ptrs = (ctypes.c_uint64 * 1)(buf.data_ptr())
    fptrs = (ctypes.c_uint64 * 1)(fbuf.data_ptr())

    def desc(rows=B, with_fi=True, row_stride=stride):
        d = _lib.PeerGatherDesc()
        d.n_peers = 1
        d.durations_ptrs = ctypes.cast(ptrs, ctypes.POINTER(ctypes.c_uint64))
        d.row0, d.rows, d.row_stride = row0, rows, row_stride
        d.frame_idx_ptrs = ctypes.cast(fptrs, ctypes.POINTER(ctypes.c_uint64)) if with_fi else None
        d.frame_idx_stride = fstride if with_fi else 0
        return d

    path, dur, fi = fused(mu_x, y, x_len, y_len, cuda, return_frame_idx=True, peer=desc())
    assert torch.equal(buf[row0:row0 + B, :T_x], dur)
    assert torch.equal(fbuf[row0:row0 + B, :T_y], fi)
    assert (buf[:row0] == -7).all() and (buf[row0 + B:] == -7).all() and (buf[:, T_x:] == -7).all()
    assert (fbuf[:row0] == -7).all() and (fbuf[row0 + B:] == -7).all() and (fbuf[:, T_y:] == -7).all()
    with pytest.raises(ValueError):           # the CUDA-core engine does not write peer memory
        fused(mu_x, y, x_len, y_len, cuda, flags=ENGINES["cuda_core"], peer=desc())
    with pytest.raises(ValueError):           # one utterance more than the peers' buffers hold
        fused(mu_x, y, x_len, y_len, cuda, peer=desc(rows=B - 1))
    with pytest.raises(ValueError):           # rows shorter than the call's T_x
        fused(mu_x, y, x_len, y_len, cuda, peer=desc(row_stride=T_x - 1))
    buf.fill_(-7)
    fbuf.fill_(-7)
    fused(mu_x, y, x_len, y_len, cuda, peer=desc(with_fi=False))
    assert torch.equal(buf[row0:row0 + B, :T_x], dur) and (fbuf == -7).all()
    buf.fill_(-7)
    fused(mu_x, y, x_len, y_len, cuda)        # no description: plain call, nothing written
    assert (buf == -7).all()


def test_path_out_and_pre_cleared_path(cuda):
    """maximum_path_from_prior(path_out=...): the path is written into the caller's tensor; with FLAG_PATH_ZEROED the
    kernel skips the clearing and only writes the 1-cells -- same result on a cleared buffer, and a dirty buffer keeps
    its dirt outside the 1-cells (proof that nothing was cleared).  Shape / dtype mismatches are refused."""
    from art_tts_b200 import _lib, monotonic_align
    rng = np.random.default_rng(12)
    B, F, T_x, T_y = 9, 80, 150, 500
    x_len = rng.integers(20, T_x + 1, B).astype(np.int32)
    y_len = np.minimum(T_y, 3 * x_len + rng.integers(0, 50, B)).astype(np.int32)
    mu_x = torch.from_numpy(rng.standard_normal((B, F, T_x)).astype(np.float32)).to(cuda)
    y = torch.from_numpy(rng.standard_normal((B, F, T_y)).astype(np.float32)).to(cuda)
    tx, ty = torch.from_numpy(x_len), torch.from_numpy(y_len)
    ref, dref = monotonic_align.maximum_path_from_prior(mu_x, None, y, tx, ty)
    out = torch.full((B, T_x, T_y), 7.0, device=cuda)
    p, d = monotonic_align.maximum_path_from_prior(mu_x, None, y, tx, ty, path_out=out)
    assert p.data_ptr() == out.data_ptr() and torch.equal(out, ref) and torch.equal(d, dref)
    out.zero_()
    p, d = monotonic_align.maximum_path_from_prior(mu_x, None, y, tx, ty, path_out=out, flags=_lib.FLAG_PATH_ZEROED)
    assert torch.equal(out, ref) and torch.equal(d, dref)
    out.fill_(7.0)
    monotonic_align.maximum_path_from_prior(mu_x, None, y, tx, ty, path_out=out, flags=_lib.FLAG_PATH_ZEROED)
    torch.cuda.synchronize()
    assert torch.equal(out[ref == 1], torch.ones_like(out[ref == 1])) and bool((out[ref == 0] == 7.0).all())
    with pytest.raises(ValueError):
        monotonic_align.maximum_path_from_prior(mu_x, None, y, tx, ty, path_out=out[:, :, :-4])
    with pytest.raises(ValueError):
        monotonic_align.maximum_path_from_prior(mu_x, None, y, tx, ty, path_out=out.to(torch.float16))
