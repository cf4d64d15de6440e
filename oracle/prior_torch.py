"""Torch-CPU restatement of the reference's log-prior + MAS block (TEST INFRASTRUCTURE).

Follows src/model/tts.py:483-505 operation by operation (same tensor ops, same order) so
that bench.py --impl reference can time "what the reference does on the host" for the fused
path: two batched matmuls + broadcast adds, then maximum_path(value, mask) through the
reference's own compiled Cython kernel (oracle/_ref) when it is present, else the C port.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import mas_oracle, ref


def log_prior_block(mu_x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """tts.py:484-495."""
    n_feats = mu_x.shape[1]
    const = -0.5 * math.log(2 * math.pi) * n_feats
    factor = -0.5 * torch.ones(mu_x.shape, dtype=mu_x.dtype, device=mu_x.device)
    y_square = torch.matmul(factor.transpose(1, 2), y ** 2)
    y_mu_double = torch.matmul(2.0 * (factor * mu_x).transpose(1, 2), y)
    mu_square = torch.sum(factor * (mu_x ** 2), 1).unsqueeze(-1)
    return y_square - y_mu_double + mu_square + const


def maximum_path_host(value: torch.Tensor, mask: torch.Tensor, kind: str = "auto", n_threads: int = 0):
    """monotonic_align/__init__.py:8-23 on CPU tensors.  kind: "omp"/"serial" = oracle/_ref
    (the reference's kernel), "port" = oracle/mas_oracle.c, "auto" = _ref omp when built."""
    value = value * mask
    dtype = value.dtype
    v = value.data.cpu().numpy().astype(np.float32)
    path = np.zeros_like(v).astype(np.int32)
    m = mask.data.cpu().numpy()
    t_x = m.sum(1)[:, 0].astype(np.int32)
    t_y = m.sum(2)[:, 0].astype(np.int32)
    if kind == "auto":
        kind = "omp" if ref.available("omp") else "port"
    if kind == "port":
        mas_oracle.maximum_path_c(path, v, t_x, t_y, n_threads=n_threads or mas_oracle.oracle_threads())
    else:
        ref.maximum_path_c(path, v, t_x, t_y, kind=kind)
    return torch.from_numpy(path).to(dtype=dtype), kind


def prior_mas_block(mu_x, y, x_mask, y_mask, kind="auto"):
    """tts.py:480-505: attn_mask, log_prior, MAS, durations."""
    attn_mask = x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)
    with torch.no_grad():
        log_prior = log_prior_block(mu_x, y)
        attn, kind = maximum_path_host(log_prior, attn_mask.squeeze(1), kind)
    dur = torch.sum(attn.unsqueeze(1), -1)
    return attn, dur, kind
