"""Shared fixtures.  `-m "not gpu"` covers the oracle against the golden vectors, the host
logic and the C-ABI surface; `-m gpu` holds the parity tests proper (CUDA path vs oracle)."""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def seeded_case(seed, B, T_x, T_y, kind):
    """Same recipe as tests/golden/make_golden.py::seeded_case (keep in sync)."""
    rng = np.random.default_rng(seed)
    if kind == "ljs":
        value = -(rng.random((B, T_x, T_y), dtype=np.float32) * 100 + 50)
        t_x = rng.integers(min(60, T_x), T_x + 1, B).astype(np.int32)
        t_y = np.minimum(T_y, 4 * t_x + rng.integers(0, 100, B)).astype(np.int32)
        t_x[0], t_y[0] = T_x, T_y
    elif kind == "full":
        value = rng.standard_normal((B, T_x, T_y), dtype=np.float32) * 5 - 100
        t_x = np.full(B, T_x, np.int32)
        t_y = np.full(B, T_y, np.int32)
    elif kind == "ties":
        value = rng.integers(-3, 1, (B, T_x, T_y)).astype(np.float32)
        t_x = rng.integers(1, T_x + 1, B).astype(np.int32)
        t_y = np.maximum(t_x, rng.integers(1, T_y + 1, B)).astype(np.int32)
        t_y = np.minimum(t_y, T_y).astype(np.int32)
        t_x = np.minimum(t_x, t_y).astype(np.int32)
    else:
        raise ValueError(kind)
    return value, t_x, t_y


def cfg3_inputs(n_vocab=149, B=64, T_x=190, T_y=872, n_feats=80, seed=3):
    """Same recipe as tests/golden/make_golden.py::cfg3_inputs (keep in sync)."""
    rng = np.random.default_rng(seed)
    x_lengths = rng.integers(60, T_x + 1, B).astype(np.int64)
    y_lengths = np.minimum(870, 4 * x_lengths + rng.integers(0, 100, B)).astype(np.int64)
    x_lengths[0], y_lengths[0] = T_x, 870
    x = rng.integers(0, n_vocab, (B, T_x)).astype(np.int64)
    y = rng.standard_normal((B, n_feats, T_y), dtype=np.float32)
    y *= (np.arange(T_y)[None, None, :] < y_lengths[:, None, None])
    return x, x_lengths, y, y_lengths


def long_text_inputs(n_vocab=149, n_feats=80, seed=44):
    """Same recipe as tests/golden/make_golden.py::long_text_inputs (keep in sync)."""
    rng = np.random.default_rng(seed)
    T_x, T_y = 420, 1320
    x_lengths = np.array([420, 257, 129, 300, 64, 385], np.int64)
    y_lengths = np.minimum(T_y, 3 * x_lengths + rng.integers(0, 61, len(x_lengths))).astype(np.int64)
    y_lengths[0] = T_y
    x = rng.integers(0, n_vocab, (len(x_lengths), T_x)).astype(np.int64)
    y = rng.standard_normal((len(x_lengths), n_feats, T_y), dtype=np.float32)
    y *= (np.arange(T_y)[None, None, :] < y_lengths[:, None, None])
    return x, x_lengths, y, y_lengths


def rect_mask(t_x, t_y, T_x, T_y, dtype=np.float32):
    m = (np.arange(T_x)[None, :, None] < np.asarray(t_x)[:, None, None]) & \
        (np.arange(T_y)[None, None, :] < np.asarray(t_y)[:, None, None])
    return m.astype(dtype)


def path_from_durations(dur, t_y, T_y):
    """Rebuild the dense path of a NORMAL (t_x <= t_y) alignment from its durations."""
    B, T_x = dur.shape
    path = np.zeros((B, T_x, T_y), np.uint8)
    for b in range(B):
        cum = np.concatenate([[0], np.cumsum(dur[b])])
        for x in range(T_x):
            path[b, x, cum[x]:min(cum[x + 1], t_y[b])] = 1
    return path


@pytest.fixture(scope="session")
def golden_small():
    return np.load(os.path.join(GOLDEN, "mas_small.npz"))


@pytest.fixture(scope="session")
def golden_seeded():
    return np.load(os.path.join(GOLDEN, "mas_seeded.npz"))


@pytest.fixture(scope="session")
def maslib():
    """The built CUDA library (compiles it here when missing/stale; nvcc needs no GPU)."""
    from art_tts_b200 import _lib, build
    build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from art_tts_b200 import build
    build.build()
    return torch.device("cuda:0")
