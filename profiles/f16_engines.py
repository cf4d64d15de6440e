import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from art_tts_b200 import monotonic_align
dev = torch.device("cuda:0")
def timeit(f, n=30, w=5):
    for _ in range(w): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (B, F, T_x, T_y) in [(1024, 16, 160, 512), (32, 16, 160, 512), (1024, 16, 190, 872)]:
    rng = np.random.default_rng(1)
    t_x = rng.integers(20, T_x + 1, B).astype(np.int32)
    t_y = np.minimum(T_y, 3 * t_x + rng.integers(0, 61, B)).astype(np.int32)
    t_x[0], t_y[0] = T_x, T_y
    o = np.argsort(-(t_x.astype(np.int64) * t_y), kind="stable")
    tx, ty = torch.from_numpy(t_x[o]).to(dev), torch.from_numpy(t_y[o]).to(dev)
    mu = torch.randn(B, F, T_x, device=dev); y = torch.randn(B, F, T_y, device=dev)
    res = []
    for flags, name in [(16, "cuda cores"), (32, "tensor cores")]:
        res.append((name, timeit(lambda: monotonic_align.maximum_path_from_prior(mu, None, y, tx, ty, flags=flags))))
    print(B, F, T_x, T_y, " | ".join(f"{n}: {t:.4f} ms" for n, t in res))
