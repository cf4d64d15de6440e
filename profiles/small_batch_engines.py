"""Configs 1-3 (one wave of utterances): fused entry through both prior engines, per-call time with the L2 flushed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from art_tts_b200 import monotonic_align, _lib
dev = torch.device("cuda", 0)
flush = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
def med(fn, iters=15):
    ts = []
    for i in range(iters + 3):
        flush.fill_(i & 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        if i >= 3: ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
for name, B, T_x, T_y, F in (("cfg1", 16, 190, 870, 80), ("cfg3", 64, 190, 872, 80), ("B=128", 128, 190, 872, 80), ("cfg2", 32, 160, 512, 16)):
    rng = np.random.default_rng(1)
    t_x = rng.integers(60, T_x + 1, B).astype(np.int32); t_y = np.minimum(min(T_y, 870), 4 * t_x + rng.integers(0, 100, B)).astype(np.int32)
    t_x[0], t_y[0] = T_x, min(T_y, 870)
    tx, ty = torch.from_numpy(t_x).to(dev), torch.from_numpy(t_y).to(dev)
    mu = torch.randn(B, F, T_x, device=dev); y = torch.randn(B, F, T_y, device=dev)
    out = [name]
    for eng, fl in (("tensor", _lib.FLAG_FORCE_TENSOR), ("cuda", _lib.FLAG_NO_TENSOR)):
        out.append(f"{eng} {med(lambda: monotonic_align.maximum_path_from_prior(mu, None, y, tx, ty, flags=fl)):.4f}")
    print("  ".join(out))
