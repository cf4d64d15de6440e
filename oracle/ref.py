"""Loader for oracle/_ref -- the reference's OWN compiled Cython kernel (test infrastructure).

oracle/build_ref.py compiles /root/reference/src/model/monotonic_align/core.pyx there; this
module imports the resulting extension by file path and wraps it the way the reference's
monotonic_align/__init__.py:8-23 does (restated here, not imported: that file cannot travel
to the GPU box and does `from .core import ...` relative to the reference tree).
"""
from __future__ import annotations

import importlib.util
import os

import numpy as np

from . import build_ref

_mods = {}


def available(kind: str = "serial") -> bool:
    return build_ref.have_ref(kind)


def load(kind: str = "serial"):
    """kind = "serial" (setup.py-faithful) or "omp" (-fopenmp -O3)."""
    if kind not in _mods:
        so = build_ref.ref_so(kind)
        if not os.path.exists(so):
            raise FileNotFoundError(f"{so} not built; run `python oracle/build_ref.py` where "
                                    "/root/reference is mounted")
        spec = importlib.util.spec_from_file_location("core", so)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _mods[kind] = mod
    return _mods[kind]


def maximum_path_c(paths, values, t_xs, t_ys, kind: str = "serial"):
    """The reference's maximum_path_c (core.pyx:40), in place on numpy arrays."""
    load(kind).maximum_path_c(paths, values, t_xs, t_ys)


def maximum_path(value, mask, kind: str = "serial"):
    """monotonic_align/__init__.py:8-23 restated around the real compiled kernel."""
    value = np.asarray(value)
    mask = np.asarray(mask)
    dtype = value.dtype
    v = np.ascontiguousarray((value * mask.astype(value.dtype, copy=False)).astype(np.float32))
    path = np.zeros(v.shape, dtype=np.int32)
    t_x = np.ascontiguousarray(mask.sum(1)[:, 0].astype(np.int32))
    t_y = np.ascontiguousarray(mask.sum(2)[:, 0].astype(np.int32))
    maximum_path_c(path, v, t_x, t_y, kind=kind)
    return path.astype(dtype)
