"""Time both entry points on the five BASELINE.json configs (resident inputs, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from art_tts_b200 import monotonic_align, _lib

dev = torch.device("cuda:0")


def timeit(f, n=10, w=3):
    for _ in range(w):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def lengths(B, T_x, T_y, kind, seed):
    rng = np.random.default_rng(seed)
    if kind == "full":
        return np.full(B, T_x, np.int32), np.full(B, T_y, np.int32)
    if kind == "ljs":
        t_x = rng.integers(60, T_x + 1, B).astype(np.int32)
        t_y = np.minimum(T_y, 4 * t_x + rng.integers(0, 100, B)).astype(np.int32)
    else:  # articulatory config 2
        t_x = rng.integers(20, T_x + 1, B).astype(np.int32)
        t_y = np.minimum(T_y, 3 * t_x + rng.integers(0, 61, B)).astype(np.int32)
    t_x[0], t_y[0] = T_x, T_y
    o = np.argsort(-(t_x.astype(np.int64) * t_y), kind="stable")
    return t_x[o], t_y[o]


cfgs = [("cfg1 B=16 190x870 F=80", 16, 80, 190, 870, "ljs"),
        ("cfg2 B=32 160x512 F=16 (articulatory)", 32, 16, 160, 512, "art"),
        ("cfg3 B=64 190x872 F=80", 64, 80, 190, 872, "ljs"),
        ("cfg4 B=32 512x4096 F=80 (long)", 32, 80, 512, 4096, "full"),
        ("cfg5 B=1024 190x872 F=80", 1024, 80, 190, 872, "ljs"),
        ("     B=1024 160x512 F=16", 1024, 16, 160, 512, "art")]
print(f"{'config':42s} {'drop-in ms':>10s} {'cells/s':>10s} {'GB/s(8B)':>9s} | {'fused ms':>9s} {'cells/s':>10s} plan")
for name, B, F, T_x, T_y, kind in cfgs:
    t_x, t_y = (torch.from_numpy(a).to(dev) for a in lengths(B, T_x, T_y, kind, 1))
    value = -(torch.rand(B, T_x, T_y, device=dev) * 100 + 50)
    mu = torch.randn(B, F, T_x, device=dev)
    y = torch.randn(B, F, T_y, device=dev)
    cells = B * T_x * T_y
    d = timeit(lambda: monotonic_align.maximum_path_lengths(value, t_x, t_y, return_durations=True))
    f = timeit(lambda: monotonic_align.maximum_path_from_prior(mu, None, y, t_x, t_y))
    lib = _lib.load()
    plan = (lib.mas_plan(B, T_x, T_y, 0), lib.mas_from_prior_plan(B, F, T_x, T_y, 0))
    print(f"{name:42s} {d:10.3f} {cells / d / 1e-3:10.3e} {8 * cells / d / 1e6:9.0f} | {f:9.3f} {cells / f / 1e-3:10.3e} {plan}")
