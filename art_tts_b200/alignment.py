"""What GradTTS/ArtTTS.compute_loss does with the alignment right after MAS,
driven by the compact outputs of the MAS kernels (durations, frame index) instead of the dense
[B,T_x,T_y] path.  Reference lines (src/model/tts.py, identical copies at :199-290, :775-870,
:1066-1160):

    logw_ / duration_loss        tts.py:503-506, model/utils.py:46-48   duration_loss_from_durations
    out_size random crop         tts.py:509-549                          crop_offsets, crop, path_segment
    mu_y = attn^T @ mu_x^T       tts.py:552-555                          aligned_mu_y (gather + segmented-sum backward)
    prior_loss                   tts.py:562-563                          aligned_mu_y_and_prior_loss
    the whole block              tts.py:477-563 minus the decoder        alignment_losses

Every tensor op here is a hand-written sm_100a kernel behind include/mas_b200.h; torch only
allocates, carries autograd edges and provides the stream.  No host synchronisation: the one
`.cpu()` of the reference's crop (tts.py:511) is avoided by drawing the offsets from the host
copy of y_lengths the data loader already has.
"""
from __future__ import annotations

import random as _random
from typing import NamedTuple, Optional, Sequence, Tuple

import torch

from . import _lib
from .monotonic_align import maximum_path_from_prior

__all__ = [
    "frame_index", "duration_targets", "duration_loss_from_durations", "crop_offsets", "crop",
    "path_segment", "aligned_mu_y", "aligned_mu_y_and_prior_loss", "alignment_losses",
    "AlignmentLosses",
]


def _i32(t, dev, n=None):
    if t is None:
        return None
    if not torch.is_tensor(t):
        t = torch.as_tensor(list(t) if not hasattr(t, "__array__") else t)
    t = t.to(device=dev, dtype=torch.int32).contiguous()
    if n is not None and t.numel() != n:
        raise ValueError(f"expected {n} entries, got {t.numel()}")
    return t


@_lib.traced
def frame_index(durations: torch.Tensor, x_lengths, y_lengths, T_y: int) -> torch.Tensor:
    """durations [B,T_x] int32 -> token index of every frame [B,T_y] int32 (-1 on padding)."""
    _lib.require_cuda(durations, "durations")
    dev = durations.device
    B, T_x = durations.shape
    d = durations.to(torch.int32).contiguous()
    out = torch.empty((B, int(T_y)), dtype=torch.int32, device=dev)
    if B == 0 or T_y == 0:
        return out
    tx, ty = _i32(x_lengths, dev, B), _i32(y_lengths, dev, B)   # keep alive across the launch
    with torch.cuda.device(dev):
        code = _lib.load().mas_frame_index(_lib.ptr(d), _lib.ptr(tx), _lib.ptr(ty), _lib.ptr(out), B,
                                           T_x, int(T_y), _lib.stream_ptr(dev))
    _lib.check(code, "mas_frame_index")
    return out


# ------------------------------------------------------------------ durations -> loss
def _duration_loss_call(logw, durations, x_lengths, want_target, want_grad, want_loss):
    dev = durations.device
    B, T_x = durations.shape
    d = durations.to(torch.int32).contiguous()
    tx = _i32(x_lengths, dev, B)
    target = torch.empty((B, T_x), dtype=torch.float32, device=dev) if want_target else None
    grad = torch.empty((B, T_x), dtype=torch.float32, device=dev) if want_grad else None
    loss = torch.empty((1,), dtype=torch.float32, device=dev) if want_loss else None
    lib = _lib.load()
    with torch.cuda.device(dev):
        ws = _lib.workspace(dev, int(lib.mas_align_workspace_bytes(max(B, 1), 1, 1))) if want_loss else None
        code = lib.mas_duration_loss_f32(_lib.ptr(logw), _lib.ptr(d), _lib.ptr(tx), _lib.ptr(target),
                                         _lib.ptr(grad), _lib.ptr(loss), B, T_x, _lib.ptr(ws),
                                         ws.numel() if ws is not None else 0, _lib.stream_ptr(dev))
    _lib.check(code, "mas_duration_loss_f32")
    return target, grad, loss


def duration_targets(durations: torch.Tensor, x_lengths) -> torch.Tensor:
    """logw_ = log(1e-8 + sum_y attn) * x_mask, shape [B,1,T_x] fp32 (tts.py:503-505)."""
    _lib.require_cuda(durations, "durations")
    target, _, _ = _duration_loss_call(None, durations, x_lengths, True, False, False)
    return target.unsqueeze(1)


class _DurationLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logw, durations, x_lengths):
        lw = logw.detach().reshape(durations.shape).to(torch.float32).contiguous()
        _, grad, loss = _duration_loss_call(lw, durations, x_lengths, False, True, True)
        ctx.save_for_backward(grad)
        ctx.shape, ctx.dtype = logw.shape, logw.dtype
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (g * grad).reshape(ctx.shape).to(ctx.dtype), None, None


@_lib.traced
def duration_loss_from_durations(logw: torch.Tensor, durations: torch.Tensor, x_lengths) -> torch.Tensor:
    """duration_loss(logw, log(1e-8 + durations) * x_mask, x_lengths) (tts.py:503-506,
    model/utils.py:46-48) as one kernel; differentiable w.r.t. logw ([B,1,T_x] or [B,T_x])."""
    _lib.require_cuda(durations, "durations")
    _lib.require_cuda(logw, "logw")
    return _DurationLoss.apply(logw, durations, x_lengths)


# ------------------------------------------------------------------ out_size crop
def crop_offsets(y_lengths_host: Sequence[int], out_size: int, rng=_random) -> Tuple[list, list]:
    """The random segment choice of tts.py:509-521, on HOST lengths (no device sync):
    offset_b = random.choice(range(0, max(y_len_b - out_size, 0))) if that range is non-empty
    else 0, drawn in batch order from `rng` (default: the `random` module, as the reference);
    seg_len_b = out_size + min(y_len_b - out_size, 0) (tts.py:539)."""
    if torch.is_tensor(y_lengths_host):
        if y_lengths_host.is_cuda:
            raise _lib.MasError("crop_offsets wants the HOST copy of y_lengths (the data loader's); "
                                "a CUDA tensor would force the sync this path removes")
        y_lengths_host = y_lengths_host.tolist()
    offsets, lens = [], []
    for n in y_lengths_host:
        n = int(n)
        end = max(n - int(out_size), 0)
        offsets.append(rng.choice(range(0, end)) if end > 0 else 0)
        lens.append(int(out_size) + min(n - int(out_size), 0))
    return offsets, lens


def crop(src: torch.Tensor, offset, seg_len, out_size: int) -> torch.Tensor:
    """dst[b,:,j] = src[b,:,offset[b]+j] for j < seg_len[b], zeros beyond (tts.py:524-544)."""
    _lib.require_cuda(src, "src")
    dev = src.device
    B, R, T_y = src.shape
    s = src.detach().to(torch.float32).contiguous()
    dst = torch.empty((B, R, int(out_size)), dtype=torch.float32, device=dev)
    off, seg = _i32(offset, dev, B), _i32(seg_len, dev, B)      # keep alive across the launch
    if dst.numel():
        with torch.cuda.device(dev):
            code = _lib.load().mas_crop_f32(_lib.ptr(s), _lib.ptr(off), _lib.ptr(seg), _lib.ptr(dst),
                                            B, R, T_y, int(out_size), _lib.stream_ptr(dev))
        _lib.check(code, "mas_crop_f32")
    return dst.to(src.dtype)


def path_segment(frame_idx: torch.Tensor, offset, seg_len, T_x: int, out_size: Optional[int] = None,
                 dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Dense attn (cropped when offset/seg_len are given) from the frame index:
    attn[b,x,j] = (frame_idx[b,offset[b]+j] == x), zeros beyond seg_len[b] (tts.py:524-544)."""
    _lib.require_cuda(frame_idx, "frame_idx")
    dev = frame_idx.device
    B, T_y = frame_idx.shape
    T_out = int(out_size) if out_size is not None else T_y
    fi = frame_idx.to(torch.int32).contiguous()
    path = torch.empty((B, int(T_x), T_out), dtype=dtype, device=dev)
    off, seg = _i32(offset, dev, B), _i32(seg_len, dev, B)      # keep alive across the launch
    if path.numel():
        with torch.cuda.device(dev):
            code = _lib.load().mas_path_segment(_lib.ptr(fi), _lib.ptr(off), _lib.ptr(seg), _lib.ptr(path),
                                                _lib.dtype_code(dtype), B, int(T_x), T_y, T_out,
                                                _lib.stream_ptr(dev))
        _lib.check(code, "mas_path_segment")
    return path


# ------------------------------------------------------------------ mu_y gather (+ prior loss)
class _AlignedGather(torch.autograd.Function):
    """(mu_x, y_seg) -> (mu_y, prior_loss); backward = segmented sum over each token's frames."""

    @staticmethod
    def forward(ctx, mu_x, y_seg, frame_idx, offset, seg_len, T_out, want_loss):
        dev = mu_x.device
        B, F, T_x = mu_x.shape
        T_y = frame_idx.shape[1]
        mu = mu_x.detach().contiguous()
        ys = y_seg.detach().contiguous() if want_loss else None
        mu_y = torch.empty((B, F, T_out), dtype=torch.float32, device=dev)
        loss = torch.zeros((2,), dtype=torch.float32, device=dev) if want_loss else None
        if B and T_out:
            lib = _lib.load()
            with torch.cuda.device(dev):
                ws = _lib.workspace(dev, int(lib.mas_align_workspace_bytes(B, F, T_out))) \
                    if want_loss else None
                code = lib.mas_align_gather_f32(
                    _lib.ptr(mu), _lib.ptr(frame_idx), _lib.ptr(offset), _lib.ptr(seg_len),
                    _lib.ptr(ys), _lib.ptr(mu_y), _lib.ptr(loss), B, F, T_x, T_y, T_out, _lib.ptr(ws),
                    ws.numel() if ws is not None else 0, _lib.stream_ptr(dev))
            _lib.check(code, "mas_align_gather_f32")
        ctx.save_for_backward(mu, ys, frame_idx, offset, seg_len, loss)
        ctx.dims = (B, F, T_x, T_y, T_out)
        ctx.want_loss = want_loss
        if want_loss:
            return mu_y, loss[0]
        return mu_y, mu_y.new_zeros(())

    @staticmethod
    def backward(ctx, g_mu_y, g_loss):
        mu, ys, frame_idx, offset, seg_len, loss = ctx.saved_tensors
        B, F, T_x, T_y, T_out = ctx.dims
        dev = mu.device
        g_mu_x = torch.zeros((B, F, T_x), dtype=torch.float32, device=dev)
        if B and T_out:
            gy = g_mu_y.to(torch.float32).contiguous() if g_mu_y is not None else None
            gl = g_loss.to(torch.float32).reshape(1).contiguous() \
                if (ctx.want_loss and g_loss is not None) else None
            norm = loss[1:2] if ctx.want_loss else None
            with torch.cuda.device(dev):
                code = _lib.load().mas_align_gather_bwd_f32(
                    _lib.ptr(gy), _lib.ptr(ys), _lib.ptr(mu), _lib.ptr(gl), _lib.ptr(norm),
                    _lib.ptr(frame_idx), _lib.ptr(offset), _lib.ptr(seg_len), _lib.ptr(g_mu_x), B, F,
                    T_x, T_y, T_out, _lib.stream_ptr(dev))
            _lib.check(code, "mas_align_gather_bwd_f32")
        return g_mu_x, None, None, None, None, None, None


def _gather(mu_x, y_seg, frame_idx, offset, seg_len, out_size, want_loss):
    _lib.require_cuda(mu_x, "mu_x")
    _lib.require_cuda(frame_idx, "frame_idx")
    dev = mu_x.device
    B = mu_x.shape[0]
    T_out = int(out_size) if out_size is not None else frame_idx.shape[1]
    fi = frame_idx.to(torch.int32).contiguous()
    mu32 = mu_x.to(torch.float32)
    ys = None
    if want_loss:
        if y_seg.shape != (B, mu_x.shape[1], T_out):
            raise ValueError(f"y_seg must be [B,F,{T_out}], got {tuple(y_seg.shape)}")
        ys = y_seg.to(torch.float32)
    mu_y, loss = _AlignedGather.apply(mu32, ys, fi, _i32(offset, dev, B), _i32(seg_len, dev, B),
                                      T_out, want_loss)
    return mu_y.to(mu_x.dtype), loss


def aligned_mu_y(mu_x: torch.Tensor, frame_idx: torch.Tensor, offset=None, seg_len=None,
                 out_size: Optional[int] = None) -> torch.Tensor:
    """mu_y [B,F,T_out] = (attn^T @ mu_x^T)^T of tts.py:552-555 as a gather through the frame
    index (bit-identical values; differentiable w.r.t. mu_x).  seg_len = frames per utterance
    (y_lengths, or the cut lengths with offset/out_size); columns beyond are zero."""
    return _gather(mu_x, None, frame_idx, offset, seg_len, out_size, False)[0]


def aligned_mu_y_and_prior_loss(mu_x: torch.Tensor, y_seg: torch.Tensor, frame_idx: torch.Tensor,
                                offset=None, seg_len=None, out_size: Optional[int] = None):
    """(mu_y, prior_loss) of tts.py:552-563 in one pass; both differentiable w.r.t. mu_x."""
    return _gather(mu_x, y_seg, frame_idx, offset, seg_len, out_size, True)


# ------------------------------------------------------------------ the whole block
class AlignmentLosses(NamedTuple):
    dur_loss: torch.Tensor          # tts.py:506
    prior_loss: torch.Tensor        # tts.py:562-563
    mu_y: torch.Tensor              # [B,F,T_out]   tts.py:552-555 (decoder input)
    y: torch.Tensor                 # [B,F,T_out]   y or y_cut (tts.py:548)
    y_mask: torch.Tensor            # [B,1,T_out]   tts.py:543-549
    y_lengths: torch.Tensor         # [B] int32     lengths of the (cut) segments
    durations: torch.Tensor         # [B,T_x] int32 (whole utterance, before the crop)
    frame_idx: torch.Tensor         # [B,T_y] int32
    attn: Optional[torch.Tensor]    # [B,T_x,T_out] only when return_attn=True
    out_offset: Optional[torch.Tensor]


@_lib.traced
def alignment_losses(mu_x: torch.Tensor, logw: torch.Tensor, x_lengths, y: torch.Tensor, y_lengths,
                     out_size: Optional[int] = None, *, y_lengths_host=None, out_offset=None,
                     rng=_random, return_attn: bool = False) -> AlignmentLosses:
    """tts.py:477-563 without the decoder call: masks, log-prior, MAS, duration loss, out_size
    crop, mu_y and prior loss, from (mu_x, logw) of the encoder and the batch (y, lengths).

    x_lengths / y_lengths: int tensors [B] (any device).  With `out_size`, the crop offsets are
    drawn like the reference (crop_offsets) from `y_lengths_host` (the data loader's copy;
    falls back to y_lengths if that is a CPU tensor) unless `out_offset` [B] is given."""
    _lib.require_cuda(mu_x, "mu_x")
    dev = mu_x.device
    B, F, T_x = mu_x.shape
    T_y = y.shape[2]
    tx, ty = _i32(x_lengths, dev, B), _i32(y_lengths, dev, B)
    with torch.no_grad():  # tts.py:483
        dur, fidx = maximum_path_from_prior(mu_x, None, y, tx, ty, want_path=False,
                                            return_frame_idx=True)
    dur_loss = duration_loss_from_durations(logw, dur, tx)
    offset = seg = None
    T_out, y_seg, seg_len = T_y, y, ty
    if out_size is not None:
        T_out = int(out_size)
        if out_offset is None:
            host = y_lengths_host if y_lengths_host is not None else y_lengths
            offs, _ = crop_offsets(host, T_out, rng)
            out_offset = torch.tensor(offs, dtype=torch.int32)
        offset = _i32(out_offset, dev, B)
        seg_len = torch.clamp(ty, max=T_out)          # tts.py:539
        y_seg = crop(y, offset, seg_len, T_out)
    mu_y, prior_loss = aligned_mu_y_and_prior_loss(mu_x, y_seg, fidx, offset, seg_len, T_out)
    y_mask = (torch.arange(T_out, device=dev)[None, :] < seg_len[:, None]).unsqueeze(1).to(mu_x.dtype)
    attn = path_segment(fidx, offset, seg_len, T_x, T_out, mu_x.dtype) if return_attn else None
    return AlignmentLosses(dur_loss, prior_loss, mu_y, y_seg, y_mask, seg_len, dur, fidx, attn, offset)


# ------------------------------------------------------------------ synthesis: durations -> alignment -> mu_y
class InferenceAlignment(NamedTuple):
    mu_y: torch.Tensor          # [B, F, y_max_length_]  (encoder_outputs of the reference = mu_y[:, :, :y_max_length])
    attn: torch.Tensor          # [B, T_x, y_max_length_] dense alignment, entries {0,1}
    y_mask: torch.Tensor        # [B, 1, y_max_length_]
    y_lengths: torch.Tensor     # [B] int64
    y_max_length: int


@_lib.traced
def inference_alignment(mu_x: torch.Tensor, logw: torch.Tensor, x_mask: torch.Tensor, length_scale: float = 1.0,
                        x_durations: Optional[torch.Tensor] = None) -> InferenceAlignment:
    """The alignment expansion of synthesis, tts.py:123-153 (ArtTTS.forward; the same text in GradTTS.forward):

        w = exp(logw) * x_mask  (or the given x_durations);  w_ceil = ceil(w) * length_scale
        y_lengths = clamp_min(sum(w_ceil), 1);  attn = generate_path(w_ceil, attn_mask);  mu_y = attn^T mu_x

    mu_x [B,F,T_x], logw / x_mask [B,1,T_x] on a CUDA device.  `attn` comes from generate_path_kernel (the
    reference's float cumsum semantics, also for a fractional length_scale), `mu_y` is a gather through the
    frame index instead of the one-hot matmul (bit-identical values).  Like the reference, the size of the
    result depends on the data: y_lengths.max() is read back once (one host synchronisation)."""
    from . import utils
    _lib.require_cuda(mu_x, "mu_x")
    B, F, T_x = mu_x.shape
    if x_durations is not None:
        w = x_durations.to(device=mu_x.device, dtype=x_mask.dtype).unsqueeze(1) * x_mask
    else:
        w = torch.exp(logw) * x_mask
    w_ceil = torch.ceil(w) * length_scale
    y_lengths = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()
    y_max_length = int(y_lengths.max()) if B else 0
    T_y = int(utils.fix_len_compatibility(y_max_length))
    y_mask = utils.sequence_mask(y_lengths, T_y).unsqueeze(1).to(x_mask.dtype)
    t_x = x_mask.reshape(B, -1).sum(1).to(torch.int32)
    attn = utils.generate_path_lengths(w_ceil.squeeze(1), t_x, y_lengths, T_y, out_dtype=mu_x.dtype)
    # token of every frame: from integer durations directly, else off the path (fractional length_scale)
    dur = w_ceil.squeeze(1)
    if float(length_scale).is_integer():
        fidx = frame_index(dur.to(torch.int32), t_x, y_lengths, T_y)
    else:
        fidx = torch.where(attn.sum(1) > 0, attn.argmax(1), torch.full_like(attn.argmax(1), -1)).to(torch.int32)
    mu_y = aligned_mu_y(mu_x, fidx, seg_len=y_lengths.to(torch.int32))
    return InferenceAlignment(mu_y, attn, y_mask, y_lengths, y_max_length)
