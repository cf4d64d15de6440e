"""CPU: the C-ABI library builds, loads and exports every symbol include/mas_b200.h declares;
argument validation answers with MAS_ERR_* before any launch (so no GPU is needed)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def declared_functions():
    src = open(os.path.join(ROOT, "include", "mas_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mas_[a-z0-9_]+)\s*\(", src)))


def test_header_and_exports_agree(maslib):
    from art_tts_b200 import _lib
    names = declared_functions()
    assert len(names) >= 9
    for n in names:
        assert hasattr(maslib, n), f"{n} declared in include/mas_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTS)


def test_no_torch_or_python_symbols_in_library(maslib):
    """The boundary is plain C: the library must not link libtorch / libpython."""
    import subprocess
    from art_tts_b200 import _lib
    out = subprocess.run(["ldd", _lib.lib_path()], capture_output=True, text=True).stdout
    assert "torch" not in out and "python" not in out


def test_version_strerror_plan_workspace(maslib):
    assert maslib.mas_abi_version() == 2
    assert maslib.mas_strerror(0) == b"ok"
    assert b"NULL" in maslib.mas_strerror(-1)
    assert maslib.mas_plan(16, 190, 870, 0) == 0        # LJSpeech shape: bits in shared memory
    assert maslib.mas_plan(32, 512, 4096, 0) == 1       # config 4: bits spill to the workspace
    assert maslib.mas_plan(2, 1024, 1500, 0) == 2       # beyond one warp: general kernel
    assert maslib.mas_plan(16, 190, 870, 1) == 2        # MAS_FLAG_FORCE_GENERAL
    assert maslib.mas_workspace_bytes(32, 512, 4096) >= 32 * 512 * 4096 // 8
    assert maslib.mas_workspace_bytes(0, 1, 1) > 0


def test_argument_validation_without_gpu(maslib):
    null = ctypes.c_void_p(None)
    one = ctypes.c_void_p(16)       # never dereferenced: validation fails first
    odd = ctypes.c_void_p(18)
    mp = maslib.mas_maximum_path
    assert mp(null, 0, null, one, one, one, 0, null, null, 1, 4, 8, null, 0, 0, null) == -1
    assert mp(one, 0, null, one, one, one, 0, null, null, 1, 0, 8, null, 0, 0, null) == -2
    assert mp(one, 4, null, one, one, one, 0, null, null, 1, 4, 8, null, 0, 0, null) == -3
    assert mp(one, 0, null, one, one, one, 6, null, null, 1, 4, 8, null, 0, 0, null) == -3
    assert mp(odd, 0, null, one, one, one, 0, null, null, 1, 4, 8, null, 0, 0, null) == -5
    # config-4 shape needs the workspace
    assert mp(one, 0, null, one, one, one, 0, null, null, 2, 512, 4096, null, 0, 0, null) == -4
    assert mp(one, 0, null, one, one, one, 0, null, null, 0, 4, 8, null, 0, 0, null) == 0  # B == 0
    gp = maslib.mas_generate_path
    assert gp(null, 4, null, null, one, 0, 1, 4, 8, null) == -1
    assert gp(one, 3, null, null, one, 0, 1, 4, 8, null) == -3
    fp = maslib.mas_from_prior_f32
    assert fp(null, null, one, one, one, one, 0, one, null, null, null, 1, 80, 4, 8, one, 1 << 20, 0,
              null) == -1
    assert fp(one, one, one, one, one, one, 0, one, null, null, null, 1, 80, 4, 8, one, 1 << 20, 0,
              null) == -3   # logs != NULL: the reference has no learned-variance prior
    lm = maslib.mas_lengths_from_mask
    assert lm(null, 0, 1, 4, 8, 32, 8, 1, one, one, null) == -1
    # consumers of the alignment
    assert maslib.mas_frame_index(null, one, one, one, 1, 4, 8, null) == -1
    assert maslib.mas_frame_index(one, one, one, odd, 1, 4, 8, null) == -5
    assert maslib.mas_duration_loss_f32(null, one, one, null, null, one, 1, 4, null, 0, null) == -1  # loss needs logw
    assert maslib.mas_duration_loss_f32(one, one, one, null, null, null, 1, 4, null, 0, null) == -1  # nothing to do
    assert maslib.mas_duration_loss_f32(one, one, one, null, null, one, 1, 4, null, 0, null) == -4   # loss needs scratch
    assert maslib.mas_crop_f32(one, null, null, null, 1, 80, 8, 4, null) == -1
    assert maslib.mas_crop_f32(one, null, null, one, 1, 80, 8, 0, null) == -2
    assert maslib.mas_path_segment(one, null, null, one, 6, 1, 4, 8, 8, null) == -3
    ag = maslib.mas_align_gather_f32
    assert ag(one, one, null, null, null, null, null, 1, 80, 4, 8, 8, null, 0, null) == -1
    assert ag(one, one, null, null, null, one, one, 1, 80, 4, 8, 8, null, 0, null) == -1   # loss needs y
    assert ag(one, one, null, null, one, one, one, 1, 80, 4, 8, 8, null, 0, null) == -4    # and scratch
    assert ag(one, one, null, null, null, one, null, 0, 80, 4, 8, 8, null, 0, null) == 0
    assert maslib.mas_align_workspace_bytes(1024, 80, 872) >= 1024 * 7 * 4
    bw = maslib.mas_align_gather_bwd_f32
    assert bw(one, null, null, null, null, null, null, null, one, 1, 80, 4, 8, 8, null) == -1
    assert bw(one, null, null, null, null, one, null, null, one, 1, 80, 1 << 20, 8, 8, null) == -2


def test_peer_gather_description_validated_without_gpu(maslib):
    """mas_peer_gather (fused all-gather over peer memory) travels with the call: its validation and
    the shape query answer before any launch, so neither touches a device."""
    from art_tts_b200 import _lib
    lib = maslib
    null, one = ctypes.c_void_p(None), ctypes.c_void_p(16)
    two = (ctypes.c_uint64 * 2)(0x7f0000000000, 0x7f0000100000)
    odd = (ctypes.c_uint64 * 1)(0x7f0000000002)

    def call(d, B=8, T_x=190, T_y=872):
        return lib.mas_from_prior_peer_f32(one, null, one, one, one, one, 0, one, null, null, null, B, 80,
                                           T_x, T_y, one, 1 << 20, 0, null, ctypes.byref(d))

    def desc(n=2, ptrs=two, row0=0, rows=1024, stride=190, fptrs=None, fstride=0):
        d = _lib.PeerGatherDesc()
        d.n_peers = n
        d.durations_ptrs = ctypes.cast(ptrs, ctypes.POINTER(ctypes.c_uint64)) if ptrs is not None else None
        d.row0, d.rows, d.row_stride = row0, rows, stride
        d.frame_idx_ptrs = ctypes.cast(fptrs, ctypes.POINTER(ctypes.c_uint64)) if fptrs is not None else None
        d.frame_idx_stride = fstride
        return d

    assert call(desc(ptrs=None)) == -2                 # pointers missing
    assert call(desc(n=17)) == -2                      # more ranks than one NVLink domain
    assert call(desc(row0=-1)) == -7
    assert call(desc(rows=0)) == -7                    # no room for a single utterance
    assert call(desc(rows=7)) == -7                    # B = 8 does not fit
    assert call(desc(stride=189)) == -7                # rows shorter than T_x
    assert call(desc(n=1, ptrs=odd)) == -5             # int32 rows need 4-byte alignment
    assert call(desc(fptrs=two, fstride=871)) == -7    # frame-index rows shorter than T_y
    assert call(desc(), B=0) == 0                      # valid description, empty batch: nothing to do
    assert b"peer" in lib.mas_strerror(-7)
    assert lib.mas_peer_durations_supported(1024, 80, 190, 872, 0) == 1    # tensor-core engine
    assert lib.mas_peer_durations_supported(1024, 16, 160, 512, 0) == 0    # CUDA-core engine (F < 32)
    assert lib.mas_peer_durations_supported(32, 80, 512, 4096, 0) == 1     # 257..512 tokens: the cluster kernel
    assert lib.mas_peer_durations_supported(32, 80, 513, 4096, 0) == 0     # beyond 512 tokens
    assert lib.mas_from_prior_plan(32, 80, 512, 4096, 0) == 0              # config 4 stays fused (prior never in HBM)
    assert lib.mas_from_prior_plan(32, 80, 512, 4096, 16) == 1             # MAS_FLAG_NO_TENSOR: prior to HBM + drop-in
    assert lib.mas_peer_durations_supported(1024, 80, 190, 872, 16) == 0   # MAS_FLAG_NO_TENSOR
    # sequence-mask lengths entry
    lm = lib.mas_lengths_from_seq_masks
    assert lm(null, 0, 4, 1, one, 0, 8, 1, 1, 4, 8, one, one, null) == -1
    assert lm(one, 9, 4, 1, one, 0, 8, 1, 1, 4, 8, one, one, null) == -3
    assert lm(one, 0, 4, 1, one, 0, 8, 1, 0, 4, 8, one, one, null) == 0


def test_python_surface_refuses_cpu_tensors(maslib):
    """No CPU fallback: the reference's host path is what this package replaces."""
    from art_tts_b200 import _lib, monotonic_align
    v = torch.zeros(1, 3, 5)
    with pytest.raises(_lib.MasError):
        monotonic_align.maximum_path(v, torch.ones_like(v))
    with pytest.raises(_lib.MasError):
        monotonic_align.maximum_path_from_prior(torch.zeros(1, 4, 3), None, torch.zeros(1, 4, 5),
                                                torch.ones(1, 1, 3), torch.ones(1, 1, 5))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from art_tts_b200 import _lib, build
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(build, "LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.MasError, match="no CPU fallback"):
        _lib.load()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "art_tts_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
