"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (build container only).

Needs /root/reference (read-only) and oracle/_ref (built by oracle/build_ref.py from the
reference's own core.pyx).  The reference's Python files are never copied into the repo: a
throw-away shim package under a temp dir symlinks them next to the compiled kernel so that
`from model import monotonic_align, GradTTS, ArtTTS` works exactly as in the reference.

Outputs (committed):
  mas_small.npz      explicit value/mask inputs + reference paths for edge cases
                     (ragged, ties, t_x=1, t_x=t_y, degenerate t_x>t_y, fp16/fp64 dtypes)
  mas_seeded.npz     larger seeded cases: seed recipe + sha256 of the inputs + reference
                     durations/score (the path is rebuilt from durations in the tests)
  prior_gradtts.npz  GradTTS(params_v2) random-init compute_loss: captured mu_x, y, log_prior,
                     attn durations, dur_loss, prior_loss        (F=80, tts.py:450-563)
  prior_arttts.npz   ArtTTS(params_v1) same capture              (F=16, tts.py:160-290)
  loss_block_gradtts.npz  GradTTS compute_loss WITH the out_size crop (tts.py:503-563): y_cut,
                     y_cut_mask, mu_y as handed to the decoder, dur_loss, prior_loss and the
                     autograd gradients of (dur_loss + prior_loss) w.r.t. mu_x and logw

  generate_path.npz  the reference's own model.utils.generate_path (utils.py:26-43) on the inference
                     call site's inputs (tts.py:130-147): int64 / fp32 durations, length_scale 1.0 / 1.3
  inference_arttts.npz   the alignment expansion of synthesis (tts.py:123-153) from ArtTTS.forward with the decoder
                     replaced by the identity: mu_y and attn for given / predicted durations, length_scale 1, 1.3, 2
  long_text_gradtts.npz  token axis of 257..420 (add-blank text; what BASELINE config 4 stands for) through the same
                     GradTTS(params_v2) compute_loss: mu_x, durations, path hash, prior rows, losses
  cfg3_gradtts.npz   BASELINE config 3 at its stated size: GradTTS(params_v2) compute_loss, B=64,
                     T_x<=190, T_y<=872, out_size=172 -- inputs by seeded recipe (+ sha256), reference
                     durations, losses, sha256 of log_prior and of the path

Run:  python tests/golden/make_golden.py            (everything)
      python tests/golden/make_golden.py --only loss  (just loss_block_gradtts.npz)
      python tests/golden/make_golden.py --only gp    (just generate_path.npz)
      python tests/golden/make_golden.py --only cfg3  (just cfg3_gradtts.npz)
      python tests/golden/make_golden.py --only long  (just long_text_gradtts.npz)
      python tests/golden/make_golden.py --only infer (just inference_arttts.npz)
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_SRC = "/root/reference/src"


def make_shim() -> str:
    from oracle import build_ref

    assert build_ref.build(), "oracle/_ref could not be built"
    tmp = tempfile.mkdtemp(prefix="ref_shim_")
    model = os.path.join(tmp, "model")
    os.makedirs(os.path.join(model, "monotonic_align"))
    for name in os.listdir(os.path.join(REF_SRC, "model")):
        src = os.path.join(REF_SRC, "model", name)
        if name == "monotonic_align" or name.startswith("__pycache__"):
            continue
        os.symlink(src, os.path.join(model, name))
    os.symlink(os.path.join(REF_SRC, "model", "monotonic_align", "__init__.py"),
               os.path.join(model, "monotonic_align", "__init__.py"))
    shutil.copy(build_ref.ref_so("serial"), os.path.join(model, "monotonic_align"))
    os.symlink(os.path.join(REF_SRC, "configs"), os.path.join(tmp, "configs"))
    return tmp


def seeded_case(seed, B, T_x, T_y, kind):
    """Input recipe shared with tests/conftest.py::seeded_case (keep in sync)."""
    rng = np.random.default_rng(seed)
    if kind == "ljs":
        value = -(rng.random((B, T_x, T_y), dtype=np.float32) * 100 + 50)
        t_x = rng.integers(min(60, T_x), T_x + 1, B).astype(np.int32)
        t_y = np.minimum(T_y, 4 * t_x + rng.integers(0, 100, B)).astype(np.int32)
        t_x[0], t_y[0] = T_x, T_y
    elif kind == "full":  # config 4 style, all full length
        value = rng.standard_normal((B, T_x, T_y), dtype=np.float32) * 5 - 100
        t_x = np.full(B, T_x, np.int32)
        t_y = np.full(B, T_y, np.int32)
    elif kind == "ties":  # integer-valued scores: many exact ties
        value = rng.integers(-3, 1, (B, T_x, T_y)).astype(np.float32)
        t_x = rng.integers(1, T_x + 1, B).astype(np.int32)
        t_y = np.maximum(t_x, rng.integers(1, T_y + 1, B)).astype(np.int32)
        t_y = np.minimum(t_y, T_y).astype(np.int32)
        t_x = np.minimum(t_x, t_y).astype(np.int32)
    else:
        raise ValueError(kind)
    return value, t_x, t_y


def rect_mask(t_x, t_y, T_x, T_y, dtype=np.float32):
    m = (np.arange(T_x)[None, :, None] < t_x[:, None, None]) & \
        (np.arange(T_y)[None, None, :] < t_y[:, None, None])
    return m.astype(dtype)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def loss_block(shim):
    """GradTTS.compute_loss with out_size (tts.py:450-565), decoder replaced by a recorder."""
    import importlib
    import random

    import torch
    from model import GradTTS, monotonic_align

    p2 = importlib.import_module("configs.params_v2")
    torch.manual_seed(p2.random_seed)
    n_vocab = 149
    g = GradTTS(n_vocab, p2.n_spks, p2.spk_emb_dim, p2.n_enc_channels, p2.filter_channels,
                p2.filter_channels_dp, p2.n_heads, p2.n_enc_layers, p2.enc_kernel, p2.enc_dropout,
                p2.window_size, p2.n_feats, p2.dec_dim, p2.beta_min, p2.beta_max, p2.pe_scale)
    g.eval()
    B, T_x, T_y, out_size, seed = 6, 40, 160, 64, 4242
    x_lengths = torch.tensor([40, 31, 17, 36, 9, 25])
    y_lengths = torch.tensor([160, 130, 80, 151, 40, 64])   # two items shorter/equal to out_size
    x = torch.randint(0, n_vocab, (B, T_x))
    y = torch.randn(B, p2.n_feats, T_y) * (torch.arange(T_y)[None, None, :] < y_lengths[:, None, None])
    enc, dec, mas = {}, {}, {}

    def enc_hook(m, i, o):
        o[0].retain_grad()
        o[1].retain_grad()
        enc.update(mu_x=o[0], logw=o[1], x_mask=o[2])

    def fake_decoder_loss(y_, y_mask_, mu_y_, spk=None):
        dec.update(y=y_.detach().clone(), y_mask=y_mask_.detach().clone(), mu_y=mu_y_.detach().clone())
        return torch.zeros(()), None

    real_mp = monotonic_align.maximum_path

    def spy(value, mask):
        path = real_mp(value, mask)
        mas["attn"] = path.detach().clone()
        return path

    h = g.encoder.register_forward_hook(enc_hook)
    g.decoder.compute_loss = fake_decoder_loss
    monotonic_align.maximum_path = spy
    random.seed(seed)
    dur_loss, prior_loss, _ = g.compute_loss(x, x_lengths, y, y_lengths, out_size=out_size)
    (dur_loss + prior_loss).backward()
    monotonic_align.maximum_path = real_mp
    h.remove()
    attn = mas["attn"].numpy()
    np.savez_compressed(
        os.path.join(HERE, "loss_block_gradtts.npz"),
        mu_x=enc["mu_x"].detach().numpy(), logw=enc["logw"].detach().numpy(),
        x_mask=enc["x_mask"].detach().numpy(), y=y.numpy(),
        x_lengths=x_lengths.numpy().astype(np.int32), y_lengths=y_lengths.numpy().astype(np.int32),
        out_size=np.int32(out_size), random_seed=np.int32(seed),
        durations=attn.sum(-1).astype(np.int32),
        y_cut=dec["y"].numpy(), y_cut_mask=dec["y_mask"].numpy(), mu_y=dec["mu_y"].numpy(),
        dur_loss=np.float32(dur_loss.item()), prior_loss=np.float32(prior_loss.item()),
        grad_mu_x=enc["mu_x"].grad.numpy(), grad_logw=enc["logw"].grad.numpy())


def generate_path_golden():
    """The reference's generate_path on what its inference path feeds it (tts.py:130-147):
    w_ceil = ceil(w) * length_scale (fp32, fractional when length_scale != 1), y_lengths =
    clamp_min(sum(w_ceil), 1).long(), mask = x_mask[..., None] * y_mask[:, :, None]; plus the
    integer durations MAS itself produces (the all-gather path rebuilds dense paths from those)."""
    import torch
    from model.utils import fix_len_compatibility, generate_path, sequence_mask

    out, names = {}, []
    rng = np.random.default_rng(2025)

    def call_site(name, w, x_lengths, length_scale):
        B, T_x = w.shape
        x_mask = sequence_mask(x_lengths, T_x).unsqueeze(1).float()
        w = w.unsqueeze(1) * x_mask
        w_ceil = torch.ceil(w) * length_scale
        y_lengths = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()
        T_y = fix_len_compatibility(int(y_lengths.max()))
        y_mask = sequence_mask(y_lengths, T_y).unsqueeze(1).to(x_mask.dtype)
        attn_mask = x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)
        path = generate_path(w_ceil.squeeze(1), attn_mask.squeeze(1))
        out[f"{name}.duration"] = w_ceil.squeeze(1).numpy()
        out[f"{name}.x_lengths"] = x_lengths.numpy().astype(np.int32)
        out[f"{name}.y_lengths"] = y_lengths.numpy().astype(np.int32)
        out[f"{name}.T_y"] = np.int32(T_y)
        out[f"{name}.path_packed"] = np.packbits(path.numpy().astype(np.uint8), axis=-1)
        out[f"{name}.path_dtype"] = np.array(str(path.dtype))
        names.append(name)

    x_lengths = torch.tensor([37, 12, 1, 30, 25])
    logw = torch.from_numpy(rng.standard_normal((5, 37)).astype(np.float32) * 0.7 + 1.0)
    call_site("fp32_scale1", torch.exp(logw), x_lengths, 1.0)
    call_site("fp32_scale1p3", torch.exp(logw), x_lengths, 1.3)
    call_site("fp32_scale0p77", torch.exp(logw), x_lengths, 0.77)
    xl = torch.tensor([190, 64, 128])
    call_site("fp32_long", torch.exp(torch.from_numpy(rng.standard_normal((3, 190)).astype(np.float32) * 0.5 + 1.2)),
              xl, 1.0)
    # integer durations with a mask given directly (what MAS produces; int64 like attn.sum(-1).long())
    for name, dt in (("int64", torch.int64), ("int32", torch.int32)):
        B, T_x, T_y = 4, 21, 64
        t_x = np.array([21, 9, 1, 16], np.int32)
        t_y = np.array([64, 30, 5, 63], np.int32)
        dur = np.zeros((B, T_x), np.int64)
        for b in range(B):   # a random monotonic alignment: every token >= 1 frame
            cuts = np.sort(rng.choice(np.arange(1, t_y[b]), t_x[b] - 1, replace=False)) if t_x[b] > 1 else []
            edges = np.concatenate([[0], cuts, [t_y[b]]]).astype(np.int64)
            dur[b, :t_x[b]] = np.diff(edges)
        mask = torch.from_numpy(rect_mask(t_x, t_y, T_x, T_y))
        path = generate_path(torch.from_numpy(dur).to(dt), mask)
        out[f"{name}.duration"] = dur.astype(np.int64 if dt == torch.int64 else np.int32)
        out[f"{name}.x_lengths"], out[f"{name}.y_lengths"] = t_x, t_y
        out[f"{name}.T_y"] = np.int32(T_y)
        out[f"{name}.path_packed"] = np.packbits(path.numpy().astype(np.uint8), axis=-1)
        out[f"{name}.path_dtype"] = np.array(str(path.dtype))
        names.append(name)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "generate_path.npz"), **out)


def cfg3_inputs(n_vocab=149, B=64, T_x=190, T_y=872, n_feats=80, seed=3):
    """Seeded inputs of the config-3 capture (shared with tests/conftest.py::cfg3_inputs, keep in
    sync): ragged LJSpeech-shape lengths, token ids U{0..148}, y ~ N(0,1) * mask."""
    rng = np.random.default_rng(seed)
    x_lengths = rng.integers(60, T_x + 1, B).astype(np.int64)
    y_lengths = np.minimum(870, 4 * x_lengths + rng.integers(0, 100, B)).astype(np.int64)
    x_lengths[0], y_lengths[0] = T_x, 870
    x = rng.integers(0, n_vocab, (B, T_x)).astype(np.int64)
    y = rng.standard_normal((B, n_feats, T_y), dtype=np.float32)
    y *= (np.arange(T_y)[None, None, :] < y_lengths[:, None, None])
    return x, x_lengths, y, y_lengths


def inference_golden():
    """The alignment expansion of synthesis (tts.py:123-153: durations -> y_lengths -> generate_path -> mu_y)
    as the reference's own ArtTTS.forward runs it, decoder replaced by the identity: given durations and
    predicted ones (exp(logw)), length_scale 1.0 and 1.3."""
    import importlib

    import torch
    from model import ArtTTS

    p1 = importlib.import_module("configs.params_v1")
    torch.manual_seed(p1.random_seed)
    a = ArtTTS(p1.n_ipa_feats, p1.n_spks, p1.spk_emb_dim, p1.n_enc_channels, p1.filter_channels,
               p1.filter_channels_dp, p1.n_heads, p1.n_enc_layers, p1.enc_kernel, p1.enc_dropout,
               p1.window_size, p1.n_feats, p1.dec_dim, p1.beta_min, p1.beta_max, p1.pe_scale)
    a.eval()
    B, T_x = 5, 48
    rng = np.random.default_rng(11)
    x_lengths = torch.tensor([48, 20, 35, 41, 1])
    x = torch.randint(-1, 2, (B, p1.n_ipa_feats, T_x)).float()
    enc = {}
    h = a.encoder.register_forward_hook(lambda m, i, o: enc.update(mu_x=o[0], logw=o[1], x_mask=o[2]))
    a.decoder.forward = lambda z, mask, mu, n_timesteps, stoc=False, spk=None: z
    out = {"x_lengths": x_lengths.numpy().astype(np.int32)}
    durs = torch.from_numpy(rng.integers(1, 9, (B, T_x)).astype(np.float32))
    cases = {"given_ls1": (durs, 1.0), "given_ls13": (durs, 1.3), "pred_ls1": (None, 1.0), "pred_ls2": (None, 2.0)}
    for name, (d, ls) in cases.items():
        enc_out, dec_out, attn = a.forward(x, x_lengths, n_timesteps=1, length_scale=ls, x_durations=d)
        out[f"{name}.length_scale"] = np.float32(ls)
        out[f"{name}.mu_y"] = enc_out.numpy()                       # [B, F, y_max_length]
        out[f"{name}.attn_packed"] = np.packbits(attn.squeeze(1).numpy().astype(np.uint8), axis=-1)
        out[f"{name}.T_y"] = np.int32(attn.shape[-1])
    h.remove()
    out["mu_x"] = enc["mu_x"].detach().numpy()
    out["logw"] = enc["logw"].detach().numpy()
    out["x_mask"] = enc["x_mask"].detach().numpy()
    out["x_durations"] = durs.numpy()
    out["names"] = np.array(list(cases))
    np.savez_compressed(os.path.join(HERE, "inference_arttts.npz"), **out)


def long_text_inputs(n_vocab=149, n_feats=80, seed=44):
    """Seeded inputs of the long-token-axis capture (shared with tests/conftest.py::long_text_inputs, keep in
    sync): add-blank phoneme text of 257..420 tokens next to short utterances, t_y = 3 t_x + U{0..60}."""
    rng = np.random.default_rng(seed)
    T_x, T_y = 420, 1320
    x_lengths = np.array([420, 257, 129, 300, 64, 385], np.int64)
    y_lengths = np.minimum(T_y, 3 * x_lengths + rng.integers(0, 61, len(x_lengths))).astype(np.int64)
    y_lengths[0] = T_y
    x = rng.integers(0, n_vocab, (len(x_lengths), T_x)).astype(np.int64)
    y = rng.standard_normal((len(x_lengths), n_feats, T_y), dtype=np.float32)
    y *= (np.arange(T_y)[None, None, :] < y_lengths[:, None, None])
    return x, x_lengths, y, y_lengths


def long_text_golden():
    """Token axis beyond 256 (what BASELINE config 4 stands for) through the reference's own compute_loss."""
    cfg3_golden(inputs=long_text_inputs, fname="long_text_gradtts.npz", recipe=[6, 420, 1320, 80, 44], row_step=7)


def cfg3_golden(inputs=None, fname="cfg3_gradtts.npz", recipe=(64, 190, 872, 80, 3), row_step=19):
    """BASELINE config 3 at its stated size (tts.py:450-563, params_v2.py:57,61)."""
    import importlib
    import random

    import torch
    from model import GradTTS, monotonic_align

    p2 = importlib.import_module("configs.params_v2")
    torch.manual_seed(p2.random_seed)
    n_vocab = 149
    g = GradTTS(n_vocab, p2.n_spks, p2.spk_emb_dim, p2.n_enc_channels, p2.filter_channels,
                p2.filter_channels_dp, p2.n_heads, p2.n_enc_layers, p2.enc_kernel, p2.enc_dropout,
                p2.window_size, p2.n_feats, p2.dec_dim, p2.beta_min, p2.beta_max, p2.pe_scale)
    g.eval()
    x, x_lengths, y, y_lengths = (inputs or cfg3_inputs)(n_vocab=n_vocab, n_feats=p2.n_feats)
    enc, dec, mas = {}, {}, {}
    h = g.encoder.register_forward_hook(lambda m, i, o: enc.update(mu_x=o[0], logw=o[1], x_mask=o[2]))

    def fake_decoder_loss(y_, y_mask_, mu_y_, spk=None):
        dec.update(y=y_.detach().clone(), y_mask=y_mask_.detach().clone(), mu_y=mu_y_.detach().clone())
        return torch.zeros(()), None

    real_mp = monotonic_align.maximum_path

    def spy(value, mask):
        mas["log_prior"] = value.detach().clone()
        path = real_mp(value, mask)
        mas["attn"] = path.detach().clone()
        return path

    g.decoder.compute_loss = fake_decoder_loss
    monotonic_align.maximum_path = spy
    seed = 373
    random.seed(seed)
    with torch.no_grad():
        dur_loss, prior_loss, _ = g.compute_loss(torch.from_numpy(x), torch.from_numpy(x_lengths),
                                                 torch.from_numpy(y), torch.from_numpy(y_lengths),
                                                 out_size=p2.out_size)
    monotonic_align.maximum_path = real_mp
    h.remove()
    attn = mas["attn"].numpy()
    lp = mas["log_prior"].numpy()
    tx, ty = x_lengths.astype(np.int32), y_lengths.astype(np.int32)
    score = np.zeros(len(tx), np.float64)   # total alignment log-likelihood = sum of lp on the path
    for b in range(len(tx)):
        score[b] = float((lp[b].astype(np.float64) * attn[b]).sum())
    np.savez_compressed(
        os.path.join(HERE, fname),
        recipe=np.array(list(recipe), np.int64), out_size=np.int32(p2.out_size),
        random_seed=np.int32(seed), x_sha256=np.array(sha(x)), y_sha256=np.array(sha(y)),
        # mu_x is an encoder output (conv/attention stack): stored, fp16-exact storage is not enough
        mu_x=enc["mu_x"].detach().numpy(), logw=enc["logw"].detach().numpy(),
        x_lengths=tx, y_lengths=ty,
        durations=attn.sum(-1).astype(np.int32), path_sha256=np.array(sha(attn.astype(np.uint8))),
        log_prior_sha256=np.array(sha(lp)), log_prior_absmax=np.float32(np.abs(lp).max()),
        # a thin sample of the prior for the 1e-5 bar without storing 42 MB: every 19th token row
        log_prior_rows=lp[:, ::row_step, :].astype(np.float32), row_step=np.int32(row_step),
        score=score, dur_loss=np.float32(dur_loss.item()), prior_loss=np.float32(prior_loss.item()),
        mu_y_sha256=np.array(sha(dec["mu_y"].numpy())), y_cut_sha256=np.array(sha(dec["y"].numpy())),
        y_cut_mask_sum=dec["y_mask"].numpy().sum(-1).astype(np.int32).reshape(-1))


def main():
    shim = make_shim()
    sys.path.insert(0, shim)
    import torch
    from model import monotonic_align  # the reference's own wrapper + compiled kernel

    if "--only" in sys.argv:
        which = sys.argv[sys.argv.index("--only") + 1]
        {"loss": lambda: loss_block(shim), "gp": generate_path_golden, "cfg3": cfg3_golden,
         "long": long_text_golden, "infer": inference_golden}[which]()
        shutil.rmtree(shim, ignore_errors=True)
        return

    def ref_path(value, mask):
        return monotonic_align.maximum_path(torch.from_numpy(value), torch.from_numpy(mask)).numpy()

    # ---------------------------------------------------------------- mas_small
    out = {}
    rng = np.random.default_rng(1234)
    cases = []
    # (name, B, T_x, T_y, t_x list or None, t_y list or None, value kind)
    v = rng.standard_normal((6, 12, 40)).astype(np.float32) * 3 - 5
    cases.append(("ragged", v, np.array([12, 7, 1, 12, 5, 9], np.int32),
                  np.array([40, 23, 9, 12, 5, 33], np.int32)))
    cases.append(("all_ties", np.zeros((1, 3, 6), np.float32), np.array([3], np.int32),
                  np.array([6], np.int32)))
    v = rng.integers(-2, 1, (5, 37, 70)).astype(np.float32)
    cases.append(("int_ties", v, np.array([37, 33, 32, 31, 2], np.int32),
                  np.array([70, 64, 33, 31, 65], np.int32)))
    v = rng.standard_normal((4, 9, 5)).astype(np.float32)
    cases.append(("degenerate_tx_gt_ty", v, np.array([9, 6, 5, 2], np.int32),
                  np.array([5, 3, 5, 1], np.int32)))
    v = rng.standard_normal((3, 70, 131)).astype(np.float32) * 10
    cases.append(("wide_unaligned", v, np.array([70, 65, 33], np.int32),
                  np.array([131, 97, 129], np.int32)))
    v = (rng.standard_normal((2, 20, 48)) * 4).astype(np.float16)
    cases.append(("fp16", v, np.array([20, 11], np.int32), np.array([48, 30], np.int32)))
    v = (rng.standard_normal((2, 20, 48)) * 4).astype(np.float64)
    cases.append(("fp64", v, np.array([20, 11], np.int32), np.array([48, 30], np.int32)))
    v = rng.standard_normal((2, 8, 16)).astype(np.float32)
    cases.append(("zero_len_frames", v, np.array([8, 3], np.int32), np.array([16, 0], np.int32)))
    names = []
    for name, value, t_x, t_y in cases:
        B, T_x, T_y = value.shape
        mask = rect_mask(t_x, t_y, T_x, T_y, value.dtype)
        # NB "zero_len_frames": an empty utterance has an all-zero mask, so the reference
        # derives t_x = t_y = 0 from it and writes nothing.
        path = ref_path(value, mask)
        assert path.dtype == value.dtype
        out[f"{name}.value"] = value
        out[f"{name}.mask"] = mask
        out[f"{name}.path"] = path
        names.append(name)
    # a non-rectangular mask (holes inside the rectangle): per-cell multiply matters
    v = rng.standard_normal((2, 10, 30)).astype(np.float32) - 1
    m = rect_mask(np.array([10, 6], np.int32), np.array([30, 17], np.int32), 10, 30)
    holes = rng.random(m.shape) < 0.15
    holes[:, :, 0] = False
    holes[:, 0, :] = False
    m = (m * ~holes).astype(np.float32)
    out["holes.value"], out["holes.mask"], out["holes.path"] = v, m, ref_path(v, m)
    names.append("holes")
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "mas_small.npz"), **out)

    # ---------------------------------------------------------------- mas_seeded
    out = {}
    recipes = [("cfg1_ljs", 0, 16, 190, 870, "ljs"),
               ("cfg1_ljs_872", 10, 8, 190, 872, "ljs"),
               ("cfg4_long", 4, 2, 512, 4096, "full"),
               ("ties_mid", 7, 12, 100, 300, "ties"),
               ("xpl_edges", 8, 10, 257, 600, "ljs"),
               ("tx_1024", 9, 2, 1024, 1500, "full")]
    for name, seed, B, T_x, T_y, kind in recipes:
        value, t_x, t_y = seeded_case(seed, B, T_x, T_y, kind)
        mask = rect_mask(t_x, t_y, T_x, T_y)
        # the reference clobbers its private copy of value; recover the score from a re-run
        from oracle import ref as oref
        vv = np.ascontiguousarray(value * mask)
        pp = np.zeros(vv.shape, np.int32)
        oref.maximum_path_c(pp, vv, t_x, t_y)
        path = ref_path(value, mask)
        assert np.array_equal(pp, path.astype(np.int32))
        score = np.array([vv[b, t_x[b] - 1, t_y[b] - 1] for b in range(B)], np.float32)
        out[f"{name}.recipe"] = np.array([seed, B, T_x, T_y], np.int64)
        out[f"{name}.kind"] = np.array(kind)
        out[f"{name}.value_sha256"] = np.array(sha(value))
        out[f"{name}.t_x"], out[f"{name}.t_y"] = t_x, t_y
        out[f"{name}.durations"] = path.sum(-1).astype(np.int32)
        out[f"{name}.score"] = score
        out[f"{name}.path_sha256"] = np.array(sha(path.astype(np.uint8)))
    out["names"] = np.array([r[0] for r in recipes])
    np.savez_compressed(os.path.join(HERE, "mas_seeded.npz"), **out)

    # ---------------------------------------------------------------- prior (full models)
    import importlib
    from model import ArtTTS, GradTTS

    captured = {}
    real_mp = monotonic_align.maximum_path

    def spy(value, mask):
        captured["log_prior"] = value.detach().clone()
        captured["mask"] = mask.detach().clone()
        path = real_mp(value, mask)
        captured["attn"] = path.detach().clone()
        return path

    def capture(model, x, x_lengths, y, y_lengths, fname, extra):
        enc_out = {}
        h = model.encoder.register_forward_hook(lambda m, i, o: enc_out.update(mu_x=o[0], logw=o[1],
                                                                                x_mask=o[2]))
        monotonic_align.maximum_path = spy
        model.eval()  # dropout off: mu_x must be reproducible from the hook only
        torch.manual_seed(99)
        with torch.no_grad():
            dur_loss, prior_loss, diff_loss = model.compute_loss(x, x_lengths, y, y_lengths,
                                                                 out_size=None)
        monotonic_align.maximum_path = real_mp
        h.remove()
        attn = captured["attn"].numpy()
        np.savez_compressed(
            os.path.join(HERE, fname),
            mu_x=enc_out["mu_x"].detach().numpy(), logw=enc_out["logw"].detach().numpy(),
            x_mask=enc_out["x_mask"].detach().numpy(), y=y.numpy(),
            x_lengths=x_lengths.numpy().astype(np.int32), y_lengths=y_lengths.numpy().astype(np.int32),
            log_prior=captured["log_prior"].numpy(), durations=attn.sum(-1).astype(np.int32),
            attn_packed=np.packbits(attn.astype(np.uint8), axis=-1),
            dur_loss=np.float32(dur_loss.item()), prior_loss=np.float32(prior_loss.item()), **extra)

    p2 = importlib.import_module("configs.params_v2")
    torch.manual_seed(p2.random_seed)
    n_vocab = 149
    g = GradTTS(n_vocab, p2.n_spks, p2.spk_emb_dim, p2.n_enc_channels, p2.filter_channels,
                p2.filter_channels_dp, p2.n_heads, p2.n_enc_layers, p2.enc_kernel, p2.enc_dropout,
                p2.window_size, p2.n_feats, p2.dec_dim, p2.beta_min, p2.beta_max, p2.pe_scale)
    B, T_x, T_y = 4, 40, 160
    x_lengths = torch.tensor([40, 31, 17, 36])
    y_lengths = torch.tensor([160, 130, 80, 151])
    x = torch.randint(0, n_vocab, (B, T_x))
    y = torch.randn(B, p2.n_feats, T_y) * (torch.arange(T_y)[None, None, :] < y_lengths[:, None, None])
    capture(g, x, x_lengths, y, y_lengths, "prior_gradtts.npz", {"x_tokens": x.numpy()})

    p1 = importlib.import_module("configs.params_v1")
    torch.manual_seed(p1.random_seed)
    a = ArtTTS(p1.n_ipa_feats, p1.n_spks, p1.spk_emb_dim, p1.n_enc_channels, p1.filter_channels,
               p1.filter_channels_dp, p1.n_heads, p1.n_enc_layers, p1.enc_kernel, p1.enc_dropout,
               p1.window_size, p1.n_feats, p1.dec_dim, p1.beta_min, p1.beta_max, p1.pe_scale)
    B, T_x, T_y = 4, 48, 152
    x_lengths = torch.tensor([48, 20, 35, 41])
    y_lengths = torch.tensor([152, 70, 121, 100])
    x = torch.randint(-1, 2, (B, p1.n_ipa_feats, T_x)).float()  # ternary IPA traits
    y = torch.randn(B, p1.n_feats, T_y)
    y[:, [12, 14]] = 0.0  # the two padded SPARC channels (data_phnm.py:136-140)
    y = y * (torch.arange(T_y)[None, None, :] < y_lengths[:, None, None])
    capture(a, x, x_lengths, y, y_lengths, "prior_arttts.npz", {"x_traits": x.numpy()})

    loss_block(shim)
    generate_path_golden()
    cfg3_golden()
    long_text_golden()
    inference_golden()
    shutil.rmtree(shim, ignore_errors=True)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
