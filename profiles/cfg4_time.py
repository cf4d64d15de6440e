"""Config 4 (B=32, 512 x 4096, F=80) per-call device time: fused entry on clusters of 4 / 2 CTAs, the
CUDA-core plan (prior to HBM + drop-in) and the drop-in kernel alone.  L2 flushed before every call."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from art_tts_b200 import monotonic_align, _lib
dev = torch.device("cuda", 0)
B, F, T_x, T_y = 32, 80, 512, 4096
g = torch.Generator(device=dev).manual_seed(4)
mu_x = torch.randn(B, F, T_x, device=dev, generator=g)
y = torch.randn(B, F, T_y, device=dev, generator=g)
value = -(torch.rand(B, T_x, T_y, device=dev, generator=g) * 100 + 50)
tx = torch.full((B,), T_x, dtype=torch.int32, device=dev)
ty = torch.full((B,), T_y, dtype=torch.int32, device=dev)
flush = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
def med(fn, iters=10):
    ts = []
    for i in range(iters + 3):
        flush.fill_(i & 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        if i >= 3: ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
for name, fl in (("cluster4", 0), ("cluster2", _lib.FLAG_CLUSTER2), ("cuda-core plan (prior to HBM)", _lib.FLAG_NO_TENSOR)):
    print(name, "fused ms:", med(lambda: monotonic_align.maximum_path_from_prior(mu_x, None, y, tx, ty, flags=fl)),
          " no dense path:", med(lambda: monotonic_align.maximum_path_from_prior(mu_x, None, y, tx, ty, flags=fl, want_path=False)))
print("drop-in ms:", med(lambda: monotonic_align.maximum_path_lengths(value, tx, ty, return_durations=True)))
