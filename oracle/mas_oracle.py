"""numpy / ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Each function cites the reference lines it restates (paths relative to /root/reference/).
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmas_oracle.so")
_lib = None

NEG = np.float32(-1e9)


def build_oracle(force: bool = False) -> str:
    """Compile mas_oracle.c -> libmas_oracle.so with the committed Makefile."""
    src = os.path.join(_HERE, "mas_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", _HERE, "-B" if force else "all"] if force
                       else ["make", "-s", "-C", _HERE], check=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build_oracle()
        lib = ctypes.CDLL(_SO)
        i32p = ctypes.POINTER(ctypes.c_int32)
        f32p = ctypes.POINTER(ctypes.c_float)
        f64p = ctypes.POINTER(ctypes.c_double)
        ci = ctypes.c_int
        lib.mas_oracle_maximum_path_c.argtypes = [i32p, f32p, i32p, i32p, ci, ci, ci,
                                                  ctypes.c_float, ci]
        lib.mas_oracle_maximum_path_c.restype = None
        lib.mas_oracle_apply_mask.argtypes = [f32p, f32p, i32p, i32p, ci, ci, ci]
        lib.mas_oracle_apply_mask.restype = None
        lib.mas_oracle_log_prior_f32.argtypes = [f32p, f32p, f32p, ci, ci, ci, ci]
        lib.mas_oracle_log_prior_f32.restype = None
        lib.mas_oracle_log_prior_f64.argtypes = [f64p, f32p, f32p, ci, ci, ci, ci]
        lib.mas_oracle_log_prior_f64.restype = None
        lib.mas_oracle_durations.argtypes = [i32p, i32p, ci, ci, ci]
        lib.mas_oracle_durations.restype = None
        lib.mas_oracle_generate_path.argtypes = [i32p, i32p, i32p, i32p, ci, ci, ci]
        lib.mas_oracle_generate_path.restype = None
        lib.mas_oracle_max_threads.argtypes = []
        lib.mas_oracle_max_threads.restype = ci
        _lib = lib
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def oracle_threads() -> int:
    return int(_load().mas_oracle_max_threads())


def sequence_mask(length, max_length=None):
    """src/model/utils.py:6-10."""
    length = np.asarray(length)
    if max_length is None:
        max_length = int(length.max())
    return np.arange(int(max_length))[None, :] < length[:, None]


def maximum_path_c(paths, values, t_xs, t_ys, max_neg_val=-1e9, n_threads=1):
    """core.pyx:38-45.  In place: `values` becomes the cumulative score table, `paths`
    (pre-zeroed int32) receives the path.  Same contiguity/dtype demands as the memoryviews."""
    assert paths.dtype == np.int32 and values.dtype == np.float32
    assert t_xs.dtype == np.int32 and t_ys.dtype == np.int32
    assert paths.flags.c_contiguous and values.flags.c_contiguous
    B, T_x, T_y = values.shape
    _load().mas_oracle_maximum_path_c(_p(paths, ctypes.c_int32), _p(values, ctypes.c_float),
                                      _p(t_xs, ctypes.c_int32), _p(t_ys, ctypes.c_int32),
                                      B, T_x, T_y, ctypes.c_float(max_neg_val), int(n_threads))


def maximum_path(value, mask, n_threads=1, return_scores=False):
    """monotonic_align/__init__.py:8-23 on numpy arrays.

    value, mask: [B, T_x, T_y].  Returns path [B, T_x, T_y] in value.dtype with {0,1}."""
    value = np.asarray(value)
    mask = np.asarray(mask)
    dtype = value.dtype
    v = (value * mask.astype(value.dtype, copy=False)).astype(np.float32)
    v = np.ascontiguousarray(v)
    path = np.zeros(v.shape, dtype=np.int32)
    t_x = mask.sum(1)[:, 0].astype(np.int32)
    t_y = mask.sum(2)[:, 0].astype(np.int32)
    maximum_path_c(path, v, np.ascontiguousarray(t_x), np.ascontiguousarray(t_y),
                   n_threads=n_threads)
    out = path.astype(dtype)
    if return_scores:
        B = v.shape[0]
        sc = np.array([v[b, t_x[b] - 1, t_y[b] - 1] if (t_x[b] >= 1 and t_y[b] >= 1) else 0.0
                       for b in range(B)], dtype=np.float32)
        return out, sc
    return out


def maximum_path_rowsweep(value, t_x, t_y):
    """One utterance, pure numpy, in the form the CUDA kernels use:
    keep one fp32 column of scores plus a 1-bit/cell direction mask
        d[x,y] = (x != 0) and (x == y or V[x,y-1] < V[x-1,y-1]),
    then backtrack with idx -= d[idx,y].  Valid for 1 <= t_x <= t_y.
    Returns (path int32 [T_x,T_y], durations int32 [T_x], score fp32)."""
    value = np.asarray(value, dtype=np.float32)
    T_x, T_y = value.shape
    assert 1 <= t_x <= t_y
    V = np.full(T_x, NEG, dtype=np.float32)
    d = np.zeros((T_x, T_y), dtype=bool)
    xs = np.arange(T_x)
    for y in range(t_y):
        prev = np.empty(T_x, dtype=np.float32)
        prev[1:] = V[:-1]
        prev[0] = np.float32(0.0) if y == 0 else NEG
        take_prev = prev > V
        d[:, y] = (xs != 0) & ((xs == y) | take_prev)
        m = np.where(take_prev, prev, V)
        new = (m + value[:, y]).astype(np.float32)
        V = np.where(xs <= y, new, NEG).astype(np.float32)
    score = V[t_x - 1]
    path = np.zeros((T_x, T_y), dtype=np.int32)
    idx = t_x - 1
    for y in range(t_y - 1, -1, -1):
        path[idx, y] = 1
        if d[idx, y]:
            idx -= 1
    return path, path.sum(1).astype(np.int32), score


def log_prior(mu_x, y, method="numpy"):
    """tts.py:483-495 in fp32, expanded form.  mu_x [B,F,T_x], y [B,F,T_y] -> [B,T_x,T_y].

    method="numpy": the literal sequence of tensor ops with numpy matmul (BLAS order);
    method="c":     mas_oracle_log_prior_f32 (ascending-f accumulation)."""
    mu_x = np.ascontiguousarray(mu_x, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.float32)
    B, F, T_x = mu_x.shape
    T_y = y.shape[2]
    if method == "c":
        lp = np.empty((B, T_x, T_y), dtype=np.float32)
        _load().mas_oracle_log_prior_f32(_p(lp, ctypes.c_float), _p(mu_x, ctypes.c_float),
                                         _p(y, ctypes.c_float), B, F, T_x, T_y)
        return lp
    const = np.float32(-0.5 * math.log(2 * math.pi) * F)
    factor = np.full(mu_x.shape, -0.5, dtype=np.float32)
    y_square = np.matmul(factor.transpose(0, 2, 1), y ** 2)
    y_mu_double = np.matmul((np.float32(2.0) * (factor * mu_x)).transpose(0, 2, 1), y)
    mu_square = np.sum(factor * (mu_x ** 2), 1)[:, :, None]
    return (y_square - y_mu_double + mu_square + const).astype(np.float32)


def log_prior_f64(mu_x, y):
    """Direct form -0.5*sum_f (y-mu)^2 - 0.5*F*log(2*pi) in fp64 (tie-breaker reference)."""
    mu_x = np.ascontiguousarray(mu_x, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.float32)
    B, F, T_x = mu_x.shape
    T_y = y.shape[2]
    lp = np.empty((B, T_x, T_y), dtype=np.float64)
    _load().mas_oracle_log_prior_f64(_p(lp, ctypes.c_double), _p(mu_x, ctypes.c_float),
                                     _p(y, ctypes.c_float), B, F, T_x, T_y)
    return lp


def durations_from_path(path):
    """tts.py:503-505: sum over the frame axis."""
    return np.asarray(path).sum(-1)


def generate_path(duration, mask):
    """src/model/utils.py:26-43 on numpy arrays (same op sequence)."""
    duration = np.asarray(duration)
    mask = np.asarray(mask)
    b, t_x, t_y = mask.shape
    # torch.cumsum on the host accumulates float32 in double (ATen's CPU accumulation type) and rounds every
    # partial sum to float32; integer durations are exact either way
    cum = (np.cumsum(duration.astype(np.float64), 1).astype(duration.dtype)
           if np.issubdtype(duration.dtype, np.floating) else np.cumsum(duration, 1))
    path = sequence_mask(cum.reshape(b * t_x), t_y).astype(mask.dtype).reshape(b, t_x, t_y)
    path = path - np.pad(path, ((0, 0), (1, 0), (0, 0)))[:, :-1]
    return path * mask


def inference_alignment(mu_x, logw, x_mask, length_scale=1.0, x_durations=None):
    """tts.py:123-153 on numpy arrays (same op sequence, fp32): returns (mu_y [B,F,T_y_], attn [B,T_x,T_y_],
    y_lengths, y_max_length)."""
    mu_x = np.asarray(mu_x, np.float32)
    x_mask = np.asarray(x_mask, np.float32)
    if x_durations is not None:
        w = np.asarray(x_durations, np.float32)[:, None, :] * x_mask
    else:
        w = np.exp(np.asarray(logw, np.float32)) * x_mask
    w_ceil = (np.ceil(w) * np.float32(length_scale)).astype(np.float32)
    y_lengths = np.maximum(w_ceil.sum((1, 2), dtype=np.float32), 1).astype(np.int64)
    y_max_length = int(y_lengths.max())
    t_y = y_max_length
    while t_y % 4:              # fix_len_compatibility (utils.py:13-17), two U-Net downsamplings
        t_y += 1
    y_mask = sequence_mask(y_lengths, t_y).astype(np.float32)[:, None, :]
    attn_mask = x_mask[:, :, :, None] * y_mask[:, :, None, :]
    attn = generate_path(w_ceil[:, 0], attn_mask[:, 0])
    mu_y = np.matmul(attn.transpose(0, 2, 1), mu_x.transpose(0, 2, 1)).transpose(0, 2, 1)
    return mu_y, attn, y_lengths, y_max_length


# ------------------------------------------------------------------------------------------
# consumers of the alignment: numpy restatements of tts.py:503-563
# ------------------------------------------------------------------------------------------
def duration_targets(attn, x_mask):
    """tts.py:503-505: logw_ = log(1e-8 + sum(attn.unsqueeze(1), -1)) * x_mask  -> [B,1,T_x]."""
    attn = np.asarray(attn, dtype=np.float32)
    s = attn.sum(-1, dtype=np.float32)[:, None, :]
    return (np.log(np.float32(1e-8) + s).astype(np.float32) * np.asarray(x_mask, np.float32))


def duration_loss(logw, logw_, lengths):
    """model/utils.py:46-48 (fp64 accumulation: the checker, not a bit-pattern)."""
    d = np.asarray(logw, np.float32) - np.asarray(logw_, np.float32)
    return np.float32((d * d).astype(np.float32).sum(dtype=np.float64) / np.asarray(lengths).sum())


def crop_offsets(y_lengths, out_size, rng):
    """tts.py:509-521: max_offset = (y_lengths - out_size).clamp(0);
    offset = random.choice(range(0, max_offset)) if max_offset > 0 else 0, in batch order."""
    out = []
    for n in np.asarray(y_lengths).tolist():
        end = max(int(n) - int(out_size), 0)
        out.append(rng.choice(range(0, end)) if end > 0 else 0)
    return np.asarray(out, dtype=np.int64)


def crop_segments(y, attn, y_lengths, out_size, out_offset):
    """tts.py:523-549: the per-item slicing loop.  Returns (y_cut, attn_cut, y_cut_lengths)."""
    y, attn = np.asarray(y), np.asarray(attn)
    B, F = y.shape[:2]
    attn_cut = np.zeros((B, attn.shape[1], out_size), attn.dtype)
    y_cut = np.zeros((B, F, out_size), y.dtype)
    lens = []
    for i in range(B):
        n = int(out_size + min(int(y_lengths[i]) - out_size, 0))
        lo = int(out_offset[i])
        y_cut[i, :, :n] = y[i, :, lo:lo + n]
        attn_cut[i, :, :n] = attn[i, :, lo:lo + n]
        lens.append(n)
    return y_cut, attn_cut, np.asarray(lens, np.int64)


def align_mu_y(attn, mu_x):
    """tts.py:552-555: mu_y = (attn^T @ mu_x^T)^T -> [B,F,T_out]."""
    attn = np.asarray(attn, np.float32)
    mu_x = np.asarray(mu_x, np.float32)
    return np.matmul(attn.transpose(0, 2, 1), mu_x.transpose(0, 2, 1)).transpose(0, 2, 1)


def prior_loss(y, mu_y, y_mask, n_feats):
    """tts.py:562-563 (elementwise in fp32, sums accumulated in fp64)."""
    y, mu_y, y_mask = (np.asarray(a, np.float32) for a in (y, mu_y, y_mask))
    e = (np.float32(0.5) * ((y - mu_y) ** 2 + np.float32(math.log(2 * math.pi)))) * y_mask
    return np.float32(e.sum(dtype=np.float64) / (y_mask.sum(dtype=np.float64) * n_feats))


def align_mu_y_grad(attn, g_mu_y):
    """Backward of align_mu_y w.r.t. mu_x: g_mu_x = (attn @ g_mu_y^T)^T -> [B,F,T_x] (fp64 sums)."""
    attn = np.asarray(attn, np.float64)
    g = np.asarray(g_mu_y, np.float64)
    return np.matmul(attn, g.transpose(0, 2, 1)).transpose(0, 2, 1)


def frame_index(path, t_y):
    """Token of every frame (argmax over the token axis), -1 beyond t_y."""
    path = np.asarray(path)
    idx = path.argmax(1).astype(np.int32)
    idx[np.arange(path.shape[2])[None, :] >= np.asarray(t_y)[:, None]] = -1
    idx[path.sum(1) == 0] = -1
    return idx
