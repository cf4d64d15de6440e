"""Host-side mirror of the alignment glue in art-tts's src/model/utils.py (same names,
argument meaning and results), backed by the sm_100a kernels where a dense [B,T_x,T_y]
tensor is produced.

    sequence_mask            src/model/utils.py:6-10     (plain torch: B*T elements)
    fix_len_compatibility    src/model/utils.py:13-17
    generate_path            src/model/utils.py:26-43, src/model_ms/utils.py:20-37  -> CUDA kernel
    duration_loss            src/model/utils.py:46-48
    mas_durations_to_logw    src/model/tts.py:503-505
"""
from __future__ import annotations

import torch

from . import _lib
from .monotonic_align import lengths_from_mask


def sequence_mask(length, max_length=None):
    if max_length is None:
        max_length = length.max()
    x = torch.arange(int(max_length), dtype=length.dtype, device=length.device)
    return x.unsqueeze(0) < length.unsqueeze(1)


def fix_len_compatibility(length, num_downsamplings_in_unet=2):
    while True:
        if length % (2 ** num_downsamplings_in_unet) == 0:
            return length
        length += 1


@_lib.traced
def generate_path_lengths(duration, t_x, t_y, T_y, out_dtype=torch.float32):
    """durations [B,T_x] (int32 or fp32) + lengths -> dense path [B,T_x,T_y] in out_dtype."""
    _lib.require_cuda(duration, "duration")
    dev = duration.device
    B, T_x = duration.shape
    if duration.dtype in (torch.int32,):
        d, code_d = duration.contiguous(), _lib.MAS_I32
    elif duration.dtype in (torch.int64, torch.int16, torch.uint8):
        d, code_d = duration.to(torch.int32).contiguous(), _lib.MAS_I32
    else:
        d, code_d = duration.to(torch.float32).contiguous(), _lib.MAS_F32
    path = torch.empty((B, T_x, int(T_y)), dtype=out_dtype, device=dev)
    if B == 0 or T_x == 0 or T_y == 0:
        return path
    t_x = None if t_x is None else t_x.to(device=dev, dtype=torch.int32).contiguous()
    t_y = None if t_y is None else t_y.to(device=dev, dtype=torch.int32).contiguous()
    lib = _lib.load()
    with torch.cuda.device(dev):
        code = lib.mas_generate_path(_lib.ptr(d), code_d, _lib.ptr(t_x), _lib.ptr(t_y),
                                     _lib.ptr(path), _lib.dtype_code(out_dtype), B, T_x, int(T_y),
                                     _lib.stream_ptr(dev))
    _lib.check(code, "mas_generate_path")
    return path


def generate_path(duration, mask):
    """Drop-in for generate_path(duration, mask) (src/model/utils.py:26-43).

    duration: [b, t_x];  mask: [b, t_x, t_y] rectangular 0/1 sequence mask.  Returns the
    path in mask.dtype.  The mask's first column/row give the lengths (as in maximum_path)."""
    b, t_x, t_y = mask.shape
    lx, ly = lengths_from_mask(mask)
    return generate_path_lengths(duration, lx, ly, t_y, out_dtype=mask.dtype)


def duration_loss(logw, logw_, lengths):
    return torch.sum((logw - logw_) ** 2) / torch.sum(lengths)


def mas_durations_to_logw(durations, x_mask):
    """logw_ = log(1e-8 + sum_y attn) * x_mask (src/model/tts.py:503-505) from int durations."""
    return torch.log(1e-8 + durations.to(x_mask.dtype).unsqueeze(1)) * x_mask
