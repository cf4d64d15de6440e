"""ctypes binding of libmas_sm100.so (the C ABI declared in include/mas_b200.h).

There is no CPU fallback: if the library is missing this module raises, loudly, at first
use.  Build it with `python -m art_tts_b200.build` (or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import build as _build

MAS_F32, MAS_F16, MAS_BF16, MAS_F64, MAS_I32, MAS_U8, MAS_I64 = range(7)
FLAG_FORCE_GENERAL = 1
FLAG_NO_ASYNC = 2
FLAG_SPILL_BITS = 4
FLAG_HOST_NO_TRIM = 8
FLAG_NO_TENSOR = 16
FLAG_FORCE_TENSOR = 32
FLAG_LOCKSTEP_DP = 64
FLAG_SKEWED_DP = 128
FLAG_TMA = 1 << 16
FLAG_CLUSTER2 = 1 << 17
FLAG_ONE_DP_WARP = 1 << 18
FLAG_PATH_ZEROED = 1 << 19
FLAG_STAGGER_MMA = 1 << 20


def flag_utt_per_cta(k: int) -> int:
    """MAS_FLAG_UTT_PER_CTA(k) of include/mas_b200.h."""
    return (int(k) & 0xff) << 8

_DTYPES = {
    torch.float32: MAS_F32,
    torch.float16: MAS_F16,
    torch.bfloat16: MAS_BF16,
    torch.float64: MAS_F64,
    torch.int32: MAS_I32,
    torch.uint8: MAS_U8,
    torch.bool: MAS_U8,
    torch.int64: MAS_I64,
}

EXPORTS = (
    "mas_abi_version", "mas_strerror", "mas_lengths_from_mask", "mas_workspace_bytes",
    "mas_maximum_path", "mas_from_prior_f32", "mas_from_prior_plan", "mas_generate_path", "mas_plan",
    "mas_launch_count",
    "mas_frame_index", "mas_duration_loss_f32", "mas_crop_f32", "mas_path_segment",
    "mas_align_workspace_bytes", "mas_align_gather_f32", "mas_align_gather_bwd_f32",
    "mas_from_prior_host_f32", "mas_set_sm_reserve",
    "mas_from_prior_peer_f32", "mas_from_prior_host_peer_f32", "mas_peer_durations_supported",
    "mas_lengths_from_seq_masks",
)


class PeerGatherDesc(ctypes.Structure):
    """`mas_peer_gather` of include/mas_b200.h: where the fused kernel stores every utterance's
    durations (and frame index) in every rank's peer-mapped buffers.  Travels with the call."""
    _fields_ = [("n_peers", ctypes.c_int),
                ("durations_ptrs", ctypes.POINTER(ctypes.c_uint64)),
                ("row0", ctypes.c_int64), ("rows", ctypes.c_int64), ("row_stride", ctypes.c_int64),
                ("frame_idx_ptrs", ctypes.POINTER(ctypes.c_uint64)),
                ("frame_idx_stride", ctypes.c_int64)]

_lib = None


class MasError(RuntimeError):
    pass


def lib_path() -> str:
    # MAS_LIB_PATH: load an instrumented build (profiles/prior_timeline.py); never a fallback
    return os.environ.get("MAS_LIB_PATH") or _build.LIB


def load() -> ctypes.CDLL:
    """Load (never build) the CUDA library; raise if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise MasError(
            f"{path} is missing: art_tts_b200 has no CPU fallback. Build the sm_100a library with "
            "`python -m art_tts_b200.build` (needs nvcc) before using it.")
    lib = ctypes.CDLL(path)
    vp, ci, i64, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t
    lib.mas_abi_version.restype = ci
    lib.mas_abi_version.argtypes = []
    lib.mas_strerror.restype = ctypes.c_char_p
    lib.mas_strerror.argtypes = [ci]
    lib.mas_lengths_from_mask.restype = ci
    lib.mas_lengths_from_mask.argtypes = [vp, ci, ci, ci, ci, i64, i64, i64, vp, vp, vp]
    lib.mas_workspace_bytes.restype = sz
    lib.mas_workspace_bytes.argtypes = [ci, ci, ci]
    lib.mas_maximum_path.restype = ci
    lib.mas_maximum_path.argtypes = [vp, ci, vp, vp, vp, vp, ci, vp, vp, ci, ci, ci, vp, sz, ci, vp]
    lib.mas_from_prior_f32.restype = ci
    lib.mas_from_prior_f32.argtypes = [vp, vp, vp, vp, vp, vp, ci, vp, vp, vp, vp, ci, ci, ci, ci,
                                       vp, sz, ci, vp]
    lib.mas_from_prior_plan.restype = ci
    lib.mas_from_prior_plan.argtypes = [ci, ci, ci, ci, ci]
    lib.mas_generate_path.restype = ci
    lib.mas_generate_path.argtypes = [vp, ci, vp, vp, vp, ci, ci, ci, ci, vp]
    lib.mas_plan.restype = ci
    lib.mas_plan.argtypes = [ci, ci, ci, ci]
    lib.mas_launch_count.restype = ctypes.c_uint64
    lib.mas_launch_count.argtypes = []
    lib.mas_frame_index.restype = ci
    lib.mas_frame_index.argtypes = [vp, vp, vp, vp, ci, ci, ci, vp]
    lib.mas_duration_loss_f32.restype = ci
    lib.mas_duration_loss_f32.argtypes = [vp, vp, vp, vp, vp, vp, ci, ci, vp, sz, vp]
    lib.mas_crop_f32.restype = ci
    lib.mas_crop_f32.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, vp]
    lib.mas_path_segment.restype = ci
    lib.mas_path_segment.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci, vp]
    lib.mas_align_workspace_bytes.restype = sz
    lib.mas_align_workspace_bytes.argtypes = [ci, ci, ci]
    lib.mas_align_gather_f32.restype = ci
    lib.mas_align_gather_f32.argtypes = [vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp, sz, vp]
    lib.mas_align_gather_bwd_f32.restype = ci
    lib.mas_align_gather_bwd_f32.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp]
    lib.mas_from_prior_host_f32.restype = ci
    lib.mas_from_prior_host_f32.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, ci, vp, vp, vp, vp, vp,
                                            ci, ci, ci, ci, vp, sz, ci, ci, vp, vp]
    lib.mas_set_sm_reserve.restype = ci
    lib.mas_set_sm_reserve.argtypes = [ci]
    pg = ctypes.POINTER(PeerGatherDesc)
    lib.mas_from_prior_peer_f32.restype = ci
    lib.mas_from_prior_peer_f32.argtypes = lib.mas_from_prior_f32.argtypes + [pg]
    lib.mas_from_prior_host_peer_f32.restype = ci
    lib.mas_from_prior_host_peer_f32.argtypes = lib.mas_from_prior_host_f32.argtypes + [pg]
    lib.mas_lengths_from_seq_masks.restype = ci
    lib.mas_lengths_from_seq_masks.argtypes = [vp, ci, i64, i64, vp, ci, i64, i64, ci, ci, ci, vp, vp, vp]
    lib.mas_peer_durations_supported.restype = ci
    lib.mas_peer_durations_supported.argtypes = [ci, ci, ci, ci, ci]
    if lib.mas_abi_version() != 2:
        raise MasError("libmas_sm100.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code == 0:
        return
    msg = load().mas_strerror(code).decode()
    if code < 0:
        raise ValueError(f"{what}: {msg} (MAS error {code})")
    raise MasError(f"{what}: CUDA error {code}: {msg}")


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPES[dt]
    except KeyError:
        raise TypeError(f"unsupported dtype {dt}") from None


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def stream_ptr(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise MasError(
            f"{name} is on {t.device}: art_tts_b200 runs MAS on a CUDA (sm_100a) device only and "
            "has no CPU fallback (the reference's host Cython path is what it replaces).")


_workspaces = {}


def workspace(device, nbytes: int) -> torch.Tensor:
    """Per (device, stream) scratch buffer, grown on demand, reused across calls."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


_NVTX = os.environ.get("MAS_NVTX", "0") not in ("", "0")


def traced(fn):
    """NVTX range around a public entry point when MAS_NVTX=1 (read at import): the ranges show up in
    Nsight Systems / `ncu --nvtx` timelines as mas.<function>; a plain passthrough otherwise."""
    if not _NVTX:
        return fn
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kw):
        torch.cuda.nvtx.range_push("mas." + fn.__name__)
        try:
            return fn(*args, **kw)
        finally:
            torch.cuda.nvtx.range_pop()
    return wrapper


def launch_count() -> int:
    return int(load().mas_launch_count())
