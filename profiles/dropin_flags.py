"""Drop-in kernel at the bench shape (B=1024, 190 x 872, ragged) under different plan flags."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from art_tts_b200 import monotonic_align, _lib
dev = torch.device("cuda", 0)
B = 1024
tx_np, ty_np = bench.make_lengths(B, 1000)
tx, ty = torch.from_numpy(tx_np).to(dev), torch.from_numpy(ty_np).to(dev)
value = -(torch.rand(B, bench.T_X, bench.T_Y, device=dev) * 100 + 50)
def t(flags, n=20):
    for _ in range(3): monotonic_align.maximum_path_lengths(value, tx, ty, return_durations=True, flags=flags)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): monotonic_align.maximum_path_lengths(value, tx, ty, return_durations=True, flags=flags)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ref = monotonic_align.maximum_path_lengths(value, tx, ty, return_durations=True)
for name, fl in (("default", 0), ("spill bits (4 CTAs/SM)", _lib.FLAG_SPILL_BITS), ("no async", _lib.FLAG_NO_ASYNC), ("tma", _lib.FLAG_TMA), ("skewed", _lib.FLAG_SKEWED_DP)):
    out = monotonic_align.maximum_path_lengths(value, tx, ty, return_durations=True, flags=fl)
    print(f"{name:26s} {t(fl):.4f} ms  same={bool(torch.equal(out[0], ref[0]))}")
print("--- components")
def t2(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
p = torch.empty(B, bench.T_X, bench.T_Y, device=dev)
print(f"durations only (band read + DP, no dense path): {t2(lambda: monotonic_align.maximum_path_lengths(value, tx, ty, return_durations=True, want_path=False)):.4f} ms")
print(f"path as uint8 (1/4 of the write bytes):          {t2(lambda: monotonic_align.maximum_path_lengths(value, tx, ty, return_durations=True, out_dtype=torch.uint8)):.4f} ms")
print(f"torch zero_ of the fp32 path alone:              {t2(lambda: p.zero_()):.4f} ms")
