"""CPU: pin the oracle (oracle/) to the reference's own outputs (tests/golden/, made by
running /root/reference in the build container) and to oracle/_ref when it is present."""
import hashlib
import os

import numpy as np
import pytest

import oracle
from oracle import ref as oref
from conftest import GOLDEN, cfg3_inputs, long_text_inputs, path_from_durations, rect_mask, seeded_case


def test_oracle_matches_reference_small_goldens(golden_small):
    g = golden_small
    for name in g["names"]:
        value, mask, want = g[f"{name}.value"], g[f"{name}.mask"], g[f"{name}.path"]
        got = oracle.maximum_path(value, mask)
        assert got.dtype == want.dtype, name
        assert np.array_equal(got, want), name


def test_oracle_matches_reference_seeded_goldens(golden_seeded):
    g = golden_seeded
    for name in g["names"]:
        seed, B, T_x, T_y = (int(v) for v in g[f"{name}.recipe"])
        value, t_x, t_y = seeded_case(seed, B, T_x, T_y, str(g[f"{name}.kind"]))
        assert hashlib.sha256(value.tobytes()).hexdigest() == str(g[f"{name}.value_sha256"]), name
        assert np.array_equal(t_x, g[f"{name}.t_x"]) and np.array_equal(t_y, g[f"{name}.t_y"])
        path, score = oracle.maximum_path(value, rect_mask(t_x, t_y, T_x, T_y), return_scores=True)
        assert np.array_equal(path.sum(-1).astype(np.int32), g[f"{name}.durations"]), name
        assert np.array_equal(score, g[f"{name}.score"]), name
        assert hashlib.sha256(path.astype(np.uint8).tobytes()).hexdigest() == \
            str(g[f"{name}.path_sha256"]), name
        # durations determine the path (generate_path is the inverse, utils.py:26-43)
        assert np.array_equal(path_from_durations(g[f"{name}.durations"], t_y, T_y),
                              path.astype(np.uint8)), name


@pytest.mark.skipif(not oref.available("serial"), reason="oracle/_ref not built on this box")
@pytest.mark.parametrize("kind", ["serial", "omp"])
def test_oracle_matches_live_reference_kernel(kind):
    rng = np.random.default_rng(42)
    for it in range(40):
        B = int(rng.integers(1, 5))
        T_x = int(rng.integers(1, 70))
        T_y = int(rng.integers(1, 150))
        t_x = rng.integers(1, T_x + 1, B).astype(np.int32)
        t_y = rng.integers(1, T_y + 1, B).astype(np.int32)   # includes degenerate t_x > t_y
        value = (rng.integers(-3, 2, (B, T_x, T_y)) if it % 2 else
                 rng.standard_normal((B, T_x, T_y)) * 4).astype(np.float32)
        mask = rect_mask(t_x, t_y, T_x, T_y)
        assert np.array_equal(oracle.maximum_path(value, mask), oref.maximum_path(value, mask, kind))


def test_oracle_threads_give_identical_paths():
    value, t_x, t_y = seeded_case(3, 12, 64, 200, "ljs")
    mask = rect_mask(t_x, t_y, 64, 200)
    assert np.array_equal(oracle.maximum_path(value, mask, n_threads=1),
                          oracle.maximum_path(value, mask, n_threads=4))


def test_rowsweep_bit_formulation_matches_oracle():
    """The 1-bit direction form the CUDA kernels use is exact."""
    rng = np.random.default_rng(7)
    for it in range(60):
        T_x = int(rng.integers(1, 50))
        T_y = int(rng.integers(T_x, 120))
        t_x = int(rng.integers(1, T_x + 1))
        t_y = int(rng.integers(t_x, T_y + 1))
        value = (rng.integers(-2, 1, (T_x, T_y)) if it % 3 == 0 else
                 rng.standard_normal((T_x, T_y))).astype(np.float32)
        mask = rect_mask([t_x], [t_y], T_x, T_y)
        want, wsc = oracle.maximum_path(value[None], mask, return_scores=True)
        got, dur, sc = oracle.maximum_path_rowsweep(value, t_x, t_y)
        assert np.array_equal(got, want[0].astype(np.int32))
        assert sc == wsc[0]
        assert dur[:t_x].min() >= 1 and dur.sum() == t_y


def test_all_ties_pile_on_last_token():
    path = oracle.maximum_path(np.zeros((1, 3, 6), np.float32), np.ones((1, 3, 6), np.float32))
    assert path.sum(-1).tolist() == [[1, 1, 4]]


def test_generate_path_is_inverse_of_mas():
    value, t_x, t_y = seeded_case(5, 6, 40, 130, "ljs")
    mask = rect_mask(t_x, t_y, 40, 130)
    path = oracle.maximum_path(value, mask)
    assert np.array_equal(oracle.generate_path(path.sum(-1), mask), path)


def test_generate_path_restatement_matches_reference_goldens():
    """oracle.generate_path against the outputs of the reference's own model.utils.generate_path
    (utils.py:26-43) on its inference call site's inputs (tts.py:130-147): fp32 durations with
    length_scale 1.0 / 1.3 / 0.77, and the int64 / int32 durations MAS produces."""
    g = np.load(os.path.join(GOLDEN, "generate_path.npz"))
    for name in g["names"]:
        dur, T_y = g[f"{name}.duration"], int(g[f"{name}.T_y"])
        mask = rect_mask(g[f"{name}.x_lengths"], g[f"{name}.y_lengths"], dur.shape[1], T_y)
        want = np.unpackbits(g[f"{name}.path_packed"], axis=-1)[:, :, :T_y]
        got = oracle.generate_path(dur, mask)
        assert got.dtype == np.float32, name              # mask.dtype, like the reference
        assert str(g[f"{name}.path_dtype"]) == "torch.float32"
        assert np.array_equal(got, want.astype(np.float32)), name


@pytest.mark.parametrize("fname,inputs", [("cfg3_gradtts.npz", cfg3_inputs), ("long_text_gradtts.npz", long_text_inputs)],
                         ids=["config3", "long_text"])
def test_config3_size_capture_against_oracle(fname, inputs):
    """BASELINE config 3 at its stated size (B=64, T_x<=190, T_y<=872, params_v2), and a token axis of 257..420
    (add-blank text: what config 4 stands for): the oracle's prior and MAS against what the reference's
    GradTTS.compute_loss produced."""
    g = np.load(os.path.join(GOLDEN, fname))
    x, x_len, y, y_len = inputs()
    assert hashlib.sha256(x.tobytes()).hexdigest() == str(g["x_sha256"])
    assert hashlib.sha256(y.tobytes()).hexdigest() == str(g["y_sha256"])
    assert np.array_equal(x_len.astype(np.int32), g["x_lengths"])
    mu_x = g["mu_x"]
    T_x, T_y = mu_x.shape[2], y.shape[2]
    st = int(g["row_step"]) if "row_step" in g.files else 19
    lp = oracle.log_prior(mu_x, y)
    rows = g["log_prior_rows"]
    sub = lp[:, ::st, :]
    m = rect_mask(x_len, y_len, T_x, T_y)[:, ::st, :].astype(bool)
    rel = np.abs(sub - rows)[m] / np.abs(rows[m])
    assert rel.max() <= 1e-5, rel.max()
    mask = rect_mask(x_len, y_len, T_x, T_y)
    path, score = oracle.maximum_path(lp, mask, return_scores=True, n_threads=4)
    want = path_from_durations(g["durations"], y_len, T_y)
    assert hashlib.sha256(want.tobytes()).hexdigest() == str(g["path_sha256"])   # durations <-> path
    assert (path.astype(np.uint8) == want).mean() >= 0.999
    assert np.allclose(score, g["score"], rtol=1e-4)
    # duration loss from the reference's durations (tts.py:503-506)
    x_mask = (np.arange(T_x)[None, None, :] < x_len[:, None, None]).astype(np.float32)
    logw_ = oracle.duration_targets(want.astype(np.float32), x_mask)
    assert np.isclose(oracle.duration_loss(g["logw"], logw_, x_len), g["dur_loss"], rtol=1e-5)


@pytest.mark.parametrize("fname,F", [("prior_gradtts.npz", 80), ("prior_arttts.npz", 16)])
def test_prior_restatement_matches_reference_model(fname, F):
    """tts.py:483-495 as captured from the reference's own compute_loss."""
    g = np.load(os.path.join(GOLDEN, fname))
    mu_x, y, want = g["mu_x"], g["y"], g["log_prior"]
    assert mu_x.shape[1] == F
    for method in ("numpy", "c"):
        got = oracle.log_prior(mu_x, y, method=method)
        assert np.allclose(got, want, rtol=1e-5, atol=1e-4), method
    direct = oracle.log_prior_f64(mu_x, y)
    assert np.allclose(direct, want, rtol=1e-5, atol=1e-4)
    # and MAS on the captured prior reproduces the captured alignment
    x_len, y_len = g["x_lengths"], g["y_lengths"]
    mask = rect_mask(x_len, y_len, mu_x.shape[2], y.shape[2])
    path = oracle.maximum_path(want, mask)
    assert np.array_equal(path.sum(-1).astype(np.int32), g["durations"])
    assert np.array_equal(np.packbits(path.astype(np.uint8), axis=-1), g["attn_packed"])


def test_loss_block_restatement_matches_reference_compute_loss():
    """tts.py:503-563 (duration targets, out_size crop, mu_y, prior loss) and the autograd
    gradients, against what the reference's own GradTTS.compute_loss produced."""
    import random
    g = np.load(os.path.join(GOLDEN, "loss_block_gradtts.npz"))
    mu_x, y, logw = g["mu_x"], g["y"], g["logw"]
    x_len, y_len, out_size = g["x_lengths"], g["y_lengths"], int(g["out_size"])
    B, F, T_x = mu_x.shape
    mask = rect_mask(x_len, y_len, T_x, y.shape[2])
    attn = oracle.maximum_path(oracle.log_prior(mu_x, y), mask)
    assert np.array_equal(attn.sum(-1).astype(np.int32), g["durations"])
    # duration loss
    logw_ = oracle.duration_targets(attn, g["x_mask"])
    assert np.isclose(oracle.duration_loss(logw, logw_, x_len), g["dur_loss"], rtol=1e-6)
    # crop: the seeded `random` stream reproduces the reference's offsets => identical y_cut
    rng = random.Random(int(g["random_seed"]))
    off = oracle.crop_offsets(y_len, out_size, rng)
    y_cut, attn_cut, cut_len = oracle.crop_segments(y, attn, y_len, out_size, off)
    assert np.array_equal(y_cut, g["y_cut"])
    y_cut_mask = oracle.sequence_mask(cut_len, out_size)[:, None, :].astype(np.float32)
    assert np.array_equal(y_cut_mask, g["y_cut_mask"])
    mu_y = oracle.align_mu_y(attn_cut, mu_x)
    assert np.array_equal(mu_y, g["mu_y"])
    assert np.isclose(oracle.prior_loss(y_cut, mu_y, y_cut_mask, F), g["prior_loss"], rtol=1e-6)
    # gradients of (dur_loss + prior_loss): d/dmu_y = (mu_y - y) * mask / (sum(mask) * F)
    g_mu_y = (mu_y - y_cut) * y_cut_mask / (y_cut_mask.sum() * F)
    assert np.allclose(oracle.align_mu_y_grad(attn_cut, g_mu_y), g["grad_mu_x"], rtol=1e-4, atol=1e-9)
    assert np.allclose(2 * (logw - logw_) / x_len.sum(), g["grad_logw"], rtol=1e-5, atol=1e-9)
    # frame index is the compact form of the path
    idx = oracle.frame_index(attn, y_len)
    for b in range(B):
        assert np.array_equal(attn[b, idx[b, :y_len[b]], np.arange(y_len[b])], np.ones(y_len[b]))
        assert (idx[b, y_len[b]:] == -1).all()


def test_inference_alignment_restatement_matches_reference_forward():
    """oracle.inference_alignment (tts.py:123-153) against the reference's own ArtTTS.forward captures."""
    g = np.load(os.path.join(GOLDEN, "inference_arttts.npz"))
    for name in g["names"]:
        given = str(name).startswith("given")
        mu_y, attn, y_len, y_max = oracle.inference_alignment(
            g["mu_x"], g["logw"], g["x_mask"], float(g[f"{name}.length_scale"]),
            g["x_durations"] if given else None)
        T_y = int(g[f"{name}.T_y"])
        want_attn = np.unpackbits(g[f"{name}.attn_packed"], axis=-1)[:, :, :T_y].astype(np.float32)
        assert attn.shape == want_attn.shape and np.array_equal(attn, want_attn), name
        assert np.array_equal(mu_y[:, :, :y_max], g[f"{name}.mu_y"]), name
