"""B200-native drop-in for art-tts's `model.monotonic_align`.

Mirrors the reference module's surface (src/model/monotonic_align/__init__.py:8-23):

    maximum_path(value, mask) -> path            same shape, dtype and device semantics

and adds the fused entry point that replaces the log-prior block of
GradTTS/ArtTTS.compute_loss (src/model/tts.py:483-505):

    maximum_path_from_prior(mu_x, logs, y, x_mask, y_mask) -> (path, durations)

Everything runs in hand-written sm_100a kernels behind the C ABI of include/mas_b200.h;
PyTorch only provides device memory and the current stream.  No host synchronisation, no
D2H/H2D copies, no CPU fallback.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from .. import _lib

__all__ = [
    "maximum_path",
    "maximum_path_lengths",
    "maximum_path_from_prior",
    "maximum_path_from_prior_host",
    "lengths_from_mask",
    "lengths_from_seq_masks",
]

_FLOAT_VALUE = (torch.float32, torch.float16, torch.bfloat16, torch.float64)


def lengths_from_mask(mask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """t_x = mask.sum(1)[:, 0], t_y = mask.sum(2)[:, 0] as int32 (reference __init__.py:18-21).

    Reads one column and one row of each utterance's mask through its strides (no copy)."""
    if mask.dim() != 3:
        raise ValueError(f"mask must be [B, T_x, T_y], got {tuple(mask.shape)}")
    _lib.require_cuda(mask, "mask")
    B, T_x, T_y = mask.shape
    t_x = torch.empty(B, dtype=torch.int32, device=mask.device)
    t_y = torch.empty(B, dtype=torch.int32, device=mask.device)
    if B == 0:
        return t_x, t_y
    lib = _lib.load()
    with torch.cuda.device(mask.device):
        sb, sx, sy = mask.stride()
        code = lib.mas_lengths_from_mask(_lib.ptr(mask), _lib.dtype_code(mask.dtype), B, T_x, T_y,
                                         sb, sx, sy, _lib.ptr(t_x), _lib.ptr(t_y),
                                         _lib.stream_ptr(mask.device))
    _lib.check(code, "mas_lengths_from_mask")
    return t_x, t_y


@_lib.traced
def maximum_path_lengths(value: torch.Tensor, t_x: torch.Tensor, t_y: torch.Tensor, *,
                         out_dtype: Optional[torch.dtype] = None,
                         cell_mask: Optional[torch.Tensor] = None,
                         return_durations: bool = False, return_score: bool = False,
                         want_path: bool = True, flags: int = 0):
    """MAS on `value [B,T_x,T_y]` with explicit int32 lengths (core.pyx:38-45 semantics).

    Returns path (dtype `out_dtype`, default value.dtype), then optionally durations
    (int32 [B,T_x]) and score (fp32 [B]) in that order."""
    if value.dim() != 3:
        raise ValueError(f"value must be [B, T_x, T_y], got {tuple(value.shape)}")
    _lib.require_cuda(value, "value")
    if value.dtype not in _FLOAT_VALUE:
        raise TypeError(f"value must be a floating tensor, got {value.dtype}")
    B, T_x, T_y = value.shape
    dev = value.device
    out_dtype = out_dtype or value.dtype
    value = value.contiguous()
    t_x = t_x.to(device=dev, dtype=torch.int32).contiguous()
    t_y = t_y.to(device=dev, dtype=torch.int32).contiguous()
    if t_x.numel() != B or t_y.numel() != B:
        raise ValueError("t_x and t_y must have one entry per utterance")
    if cell_mask is not None:
        if cell_mask.shape != value.shape:
            raise ValueError("cell_mask must have value's shape")
        cell_mask = cell_mask.to(device=dev, dtype=torch.float32).contiguous()
    path = torch.empty((B, T_x, T_y), dtype=out_dtype, device=dev) if want_path else None
    dur = torch.empty((B, T_x), dtype=torch.int32, device=dev) if return_durations else None
    score = torch.empty((B,), dtype=torch.float32, device=dev) if return_score else None
    if B > 0 and T_x > 0 and T_y > 0:
        lib = _lib.load()
        with torch.cuda.device(dev):
            nws = int(lib.mas_workspace_bytes(B, T_x, T_y))
            ws = _lib.workspace(dev, nws) if lib.mas_plan(B, T_x, T_y, flags) != 0 else None
            code = lib.mas_maximum_path(
                _lib.ptr(value), _lib.dtype_code(value.dtype), _lib.ptr(cell_mask), _lib.ptr(t_x),
                _lib.ptr(t_y), _lib.ptr(path), _lib.dtype_code(out_dtype), _lib.ptr(dur),
                _lib.ptr(score), B, T_x, T_y, _lib.ptr(ws), ws.numel() if ws is not None else 0,
                flags, _lib.stream_ptr(dev))
        _lib.check(code, "mas_maximum_path")
    out = [path] if want_path else []
    if return_durations:
        out.append(dur)
    if return_score:
        out.append(score)
    return out[0] if len(out) == 1 else tuple(out)


@_lib.traced
def maximum_path(value: torch.Tensor, mask: torch.Tensor, *, strict_mask: bool = True, flags: int = 0):
    """Drop-in for the reference's `maximum_path(value, mask)` (__init__.py:8-23).

    value: [b, t_x, t_y] float tensor on a CUDA device;  mask: [b, t_x, t_y] 0/1.
    Returns the hard monotonic alignment [b, t_x, t_y] with entries {0,1}, in the dtype of
    `value * mask` and on value's device -- bit-exact with the reference's Cython kernel.

    Like the reference, every cell is multiplied by the mask (__init__.py:13) and the lengths
    are read from the mask's first column / row (__init__.py:20-21): exact for ANY mask,
    12 B/cell of HBM traffic.  For the rectangular sequence masks the reference's callers build
    (tts.py:477-480: x_mask[..., None] * y_mask[:, :, None]) the multiplication is the identity
    on every cell the algorithm touches; pass strict_mask=False there (or call
    maximum_path_lengths) and only that column and row of the mask are read (8 B/cell).
    strict_mask=False on a mask with holes gives a different path than the reference."""
    if value.shape != mask.shape:
        raise ValueError(f"value {tuple(value.shape)} and mask {tuple(mask.shape)} differ")
    _lib.require_cuda(value, "value")
    mask = mask.to(value.device)
    out_dtype = torch.result_type(value, mask)  # dtype of `value * mask`
    if out_dtype not in _FLOAT_VALUE:
        raise TypeError(f"value * mask has dtype {out_dtype}; a floating dtype is required")
    t_x, t_y = lengths_from_mask(mask)
    cell_mask = None
    if strict_mask:
        if value.dtype == torch.float32 and mask.dtype == torch.float32:
            cell_mask = mask
        else:  # exact promotion semantics of `value * mask` for mixed dtypes
            value = value * mask
    if value.dtype not in _FLOAT_VALUE:
        value = value.to(out_dtype)
    return maximum_path_lengths(value, t_x, t_y, out_dtype=out_dtype, cell_mask=cell_mask, flags=flags)


def lengths_from_seq_masks(x_mask: torch.Tensor, y_mask: torch.Tensor,
                           T_x: int, T_y: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """t_x = sum(x_mask[b]), t_y = sum(y_mask[b]) as int32, one kernel launch: the lengths the
    reference would read off attn_mask = x_mask[..., None] * y_mask[:, :, None] (tts.py:477-480,
    __init__.py:20-21).  Masks are [B,1,T] or [B,T], any dtype, any strides."""
    _lib.require_cuda(x_mask, "x_mask")
    _lib.require_cuda(y_mask, "y_mask")
    B = x_mask.shape[0]
    xm = x_mask.reshape(B, -1) if x_mask.dim() != 2 else x_mask
    ym = y_mask.reshape(B, -1) if y_mask.dim() != 2 else y_mask
    if xm.shape[1] != T_x or ym.shape[1] != T_y or ym.shape[0] != B:
        raise ValueError("mask length does not match the tensor it masks")
    dev = x_mask.device
    t_x = torch.empty(B, dtype=torch.int32, device=dev)
    t_y = torch.empty(B, dtype=torch.int32, device=dev)
    if B == 0:
        return t_x, t_y
    with torch.cuda.device(dev):
        code = _lib.load().mas_lengths_from_seq_masks(
            _lib.ptr(xm), _lib.dtype_code(xm.dtype), xm.stride(0), xm.stride(1),
            _lib.ptr(ym), _lib.dtype_code(ym.dtype), ym.stride(0), ym.stride(1),
            B, T_x, T_y, _lib.ptr(t_x), _lib.ptr(t_y), _lib.stream_ptr(dev))
    _lib.check(code, "mas_lengths_from_seq_masks")
    return t_x, t_y


@_lib.traced
def maximum_path_from_prior(mu_x: torch.Tensor, logs: Optional[torch.Tensor], y: torch.Tensor,
                            x_mask: torch.Tensor, y_mask: torch.Tensor, *,
                            return_score: bool = False, return_frame_idx: bool = False,
                            return_log_prior: bool = False, want_path: bool = True,
                            flags: int = 0, peer=None, path_out: Optional[torch.Tensor] = None):
    """Fused Gaussian log-prior + MAS + durations (replaces tts.py:483-505).

    mu_x [B,F,T_x], y [B,F,T_y] fp32;  x_mask [B,1,T_x], y_mask [B,1,T_y] 0/1 sequence masks
    (or int lengths [B]).  `logs` must be None: the reference's prior has unit variance.
    `peer`: a `_lib.PeerGatherDesc` (see distributed.PeerDurationGather) -- the kernel also stores
    the durations rows into every rank's peer-mapped buffer.
    `path_out`: write the path into this [B,T_x,T_y] tensor of mu_x.dtype instead of a new one; with
    `flags | FLAG_PATH_ZEROED` the caller guarantees it is all zeros already (see `ZeroedPathPool`).
    Returns (path [B,T_x,T_y] in mu_x.dtype, durations [B,T_x] int32), followed by the
    optional extras in the order score [B] fp32, frame_idx [B,T_y] int32, log_prior."""
    if logs is not None:
        raise NotImplementedError("the reference has a unit-variance prior only (logs=None)")
    if mu_x.dim() != 3 or y.dim() != 3 or mu_x.shape[:2] != y.shape[:2]:
        raise ValueError("mu_x must be [B,F,T_x] and y [B,F,T_y]")
    _lib.require_cuda(mu_x, "mu_x")
    _lib.require_cuda(y, "y")
    dev = mu_x.device
    B, F, T_x = mu_x.shape
    T_y = y.shape[2]
    out_dtype = mu_x.dtype
    mu32 = mu_x.detach().to(torch.float32).contiguous()
    y32 = y.detach().to(torch.float32).contiguous()

    if x_mask.dim() == 1 and y_mask.dim() == 1:       # int lengths [B]
        t_x = x_mask.to(device=dev, dtype=torch.int32).contiguous()
        t_y = y_mask.to(device=dev, dtype=torch.int32).contiguous()
        if t_x.numel() != B or t_y.numel() != B:
            raise ValueError("lengths must have one entry per utterance")
    elif x_mask.dim() == 1 or y_mask.dim() == 1:
        raise ValueError("pass either two masks or two length vectors")
    else:                                              # sequence masks: one kernel, no eager torch ops
        t_x, t_y = lengths_from_seq_masks(x_mask.to(dev), y_mask.to(dev), T_x, T_y)
    if path_out is not None:
        if (not want_path or tuple(path_out.shape) != (B, T_x, T_y) or path_out.dtype != out_dtype
                or path_out.device != dev or not path_out.is_contiguous()):
            raise ValueError("path_out must be a contiguous [B,T_x,T_y] tensor of mu_x's dtype on mu_x's device")
        path = path_out
    else:
        flags &= ~_lib.FLAG_PATH_ZEROED
        path = torch.empty((B, T_x, T_y), dtype=out_dtype, device=dev) if want_path else None
    dur = torch.empty((B, T_x), dtype=torch.int32, device=dev)
    score = torch.empty((B,), dtype=torch.float32, device=dev) if return_score else None
    fidx = torch.empty((B, T_y), dtype=torch.int32, device=dev) if return_frame_idx else None
    lp = None
    if B > 0 and T_x > 0 and T_y > 0:
        lib = _lib.load()
        unfused = lib.mas_from_prior_plan(B, F, T_x, T_y, flags) != 0
        if return_log_prior or unfused:
            lp = torch.empty((B, T_x, T_y), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            nws = int(lib.mas_workspace_bytes(B, T_x, T_y))
            ws = _lib.workspace(dev, nws)
            code = lib.mas_from_prior_peer_f32(
                _lib.ptr(mu32), None, _lib.ptr(y32), _lib.ptr(t_x), _lib.ptr(t_y), _lib.ptr(path),
                _lib.dtype_code(out_dtype), _lib.ptr(dur), _lib.ptr(fidx), _lib.ptr(score),
                _lib.ptr(lp), B, F, T_x, T_y, _lib.ptr(ws), ws.numel(), flags, _lib.stream_ptr(dev),
                ctypes.byref(peer) if peer is not None else None)
        _lib.check(code, "mas_from_prior_f32")
    out = [path, dur] if want_path else [dur]
    for flag, t in ((return_score, score), (return_frame_idx, fidx), (return_log_prior, lp)):
        if flag:
            out.append(t)
    return tuple(out) if len(out) > 1 else out[0]


_staging = {}


def _host_staging(dev, B, F, T_x, T_y):
    """Device staging of the host-buffer entry: one grow-only flat buffer per (device, stream) -- the
    library orders its copies against the caller's stream only, so two streams must not share one --
    carved into the padded shapes of this call (varying padded shapes reuse the same memory)."""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    def up4(n):               # keep every piece 16-byte aligned (vectorised loads of the kernels)
        return (n + 3) // 4 * 4

    n1, n2 = B * F * T_x, B * F * T_y
    o1 = up4(n1)
    o2 = o1 + up4(n2)
    o3 = o2 + up4(B)
    need = o3 + B
    buf = _staging.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.empty(max(need, 1 << 16), dtype=torch.float32, device=dev)
        _staging[key] = buf
    return (buf[:n1].view(B, F, T_x), buf[o1:o1 + n2].view(B, F, T_y),
            buf[o2:o2 + B].view(torch.int32), buf[o3:o3 + B].view(torch.int32))


@_lib.traced
def maximum_path_from_prior_host(mu_x: torch.Tensor, y: torch.Tensor, x_lengths: torch.Tensor,
                                 y_lengths: torch.Tensor, device=None, *, want_path: bool = True,
                                 out_dtype: torch.dtype = torch.float32, chunk: int = 0,
                                 durations_host: Optional[torch.Tensor] = None,
                                 score_host: Optional[torch.Tensor] = None, flags: int = 0, peer=None):
    """Fused prior + MAS for a batch that still lives in HOST memory (the data loader's pinned
    tensors; train_v2.py:203 -> tts.py:466 moves the whole padded batch first).

    mu_x [B,F,T_x], y [B,F,T_y] fp32 CPU tensors (pin them: pageable memory makes the copies
    synchronous), x_lengths / y_lengths int32 CPU tensors [B].  Only the frames/tokens MAS can
    touch are copied (chunks of the length-bucketed batch trimmed to their longest utterance)
    and each chunk's kernel starts as soon as its rows have landed.  Returns
    (path or None, durations, score, h2d_bytes) on `device`; when `durations_host` /
    `score_host` (pinned) are given they are filled asynchronously on the current stream --
    synchronise it before reading them."""
    for name, t in (("mu_x", mu_x), ("y", y), ("x_lengths", x_lengths), ("y_lengths", y_lengths)):
        if t.is_cuda:
            raise _lib.MasError(f"{name} is already on {t.device}: use maximum_path_from_prior")
    if mu_x.dtype != torch.float32 or y.dtype != torch.float32:
        raise TypeError("mu_x and y must be float32")
    if mu_x.dim() != 3 or y.dim() != 3 or mu_x.shape[:2] != y.shape[:2]:
        raise ValueError("mu_x must be [B,F,T_x] and y [B,F,T_y]")
    dev = torch.device(device if device is not None else "cuda")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    B, F, T_x = mu_x.shape
    T_y = y.shape[2]
    mu_x, y = mu_x.contiguous(), y.contiguous()
    tx = x_lengths.to(torch.int32).contiguous()
    ty = y_lengths.to(torch.int32).contiguous()
    if tx.numel() != B or ty.numel() != B:
        raise ValueError("x_lengths and y_lengths must have one entry per utterance")
    lib = _lib.load()
    if B and lib.mas_from_prior_plan(B, F, T_x, T_y, flags) != 0:
        raise ValueError("shape needs the unfused plan; copy the batch and call maximum_path_from_prior")
    with torch.cuda.device(dev):
        st = _host_staging(dev, B, F, T_x, T_y)
    path = torch.empty((B, T_x, T_y), dtype=out_dtype, device=dev) if want_path else None
    dur = torch.empty((B, T_x), dtype=torch.int32, device=dev)
    score = torch.empty((B,), dtype=torch.float32, device=dev)
    moved = ctypes.c_uint64(0)
    if B:
        with torch.cuda.device(dev):
            ws = _lib.workspace(dev, int(lib.mas_workspace_bytes(B, T_x, T_y)))
            code = lib.mas_from_prior_host_peer_f32(
                _lib.ptr(mu_x), _lib.ptr(y), _lib.ptr(tx), _lib.ptr(ty), _lib.ptr(st[0]),
                _lib.ptr(st[1]), _lib.ptr(st[2]), _lib.ptr(st[3]), _lib.ptr(path),
                _lib.dtype_code(out_dtype), _lib.ptr(dur), None, _lib.ptr(score),
                _lib.ptr(durations_host), _lib.ptr(score_host), B, F, T_x, T_y, _lib.ptr(ws),
                ws.numel(), int(chunk), flags, _lib.stream_ptr(dev), ctypes.byref(moved),
                ctypes.byref(peer) if peer is not None else None)
        _lib.check(code, "mas_from_prior_host_f32")
    return path, dur, score, int(moved.value)
