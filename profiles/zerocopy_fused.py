"""e2e experiment: the fused kernel reading mu_x / y straight out of pinned host memory (zero-copy over PCIe),
against the staged host entry (trimmed chunked H2D by the copy engine, kernels behind the chunks)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from art_tts_b200 import monotonic_align, _lib
dev = torch.device("cuda", 0)
B = 1024
T_X, T_Y, F = bench.T_X, bench.T_Y, bench.N_FEATS
tx_np, ty_np = bench.make_lengths(B, 1000)
g = torch.Generator().manual_seed(1)
xm = (torch.arange(T_X)[None, :] < torch.from_numpy(tx_np)[:, None]).float()
ym = (torch.arange(T_Y)[None, :] < torch.from_numpy(ty_np)[:, None]).float()
mu_h = (torch.randn(B, F, T_X, generator=g) * xm[:, None, :]).pin_memory()
y_h = (torch.randn(B, F, T_Y, generator=g) * ym[:, None, :]).pin_memory()
tx_h, ty_h = torch.from_numpy(tx_np).pin_memory(), torch.from_numpy(ty_np).pin_memory()
h_dur = torch.empty(B, T_X, dtype=torch.int32).pin_memory()
h_score = torch.empty(B, dtype=torch.float32).pin_memory()
lib = _lib.load()
path = torch.empty(B, T_X, T_Y, device=dev)
dur = torch.empty(B, T_X, dtype=torch.int32, device=dev)
score = torch.empty(B, dtype=torch.float32, device=dev)
tx_d, ty_d = torch.empty(B, dtype=torch.int32, device=dev), torch.empty(B, dtype=torch.int32, device=dev)
ws = torch.empty(int(lib.mas_workspace_bytes(B, T_X, T_Y)), dtype=torch.uint8, device=dev)


def zero_copy_step():
    tx_d.copy_(tx_h, non_blocking=True)
    ty_d.copy_(ty_h, non_blocking=True)
    code = lib.mas_from_prior_f32(_lib.ptr(mu_h), None, _lib.ptr(y_h), _lib.ptr(tx_d), _lib.ptr(ty_d), _lib.ptr(path), 0,
                                  _lib.ptr(dur), None, _lib.ptr(score), None, B, F, T_X, T_Y, _lib.ptr(ws), ws.numel(), 0,
                                  _lib.stream_ptr(dev))
    assert code == 0, code
    h_dur.copy_(dur, non_blocking=True)
    h_score.copy_(score, non_blocking=True)
    torch.cuda.current_stream().synchronize()


def staged_step():
    monotonic_align.maximum_path_from_prior_host(mu_h, y_h, tx_h, ty_h, dev, durations_host=h_dur, score_host=h_score)
    torch.cuda.current_stream().synchronize()


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


staged_step()
ref = h_dur.clone()
zero_copy_step()
print("same durations:", bool(torch.equal(ref, h_dur)))
print(f"staged host entry : {timeit(staged_step):.3f} ms per step")
print(f"zero-copy kernel  : {timeit(zero_copy_step):.3f} ms per step")
valid = 4 * F * int(tx_np.astype(np.int64).sum() + ty_np.astype(np.int64).sum())
print(f"valid input bytes {valid / 1e6:.1f} MB")
