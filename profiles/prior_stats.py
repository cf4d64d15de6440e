"""Coarse per-CTA cycle accounting of the persistent fused kernel (MAS_PRIOR_STATS=1).
   MAS_PRIOR_STATS=1 python profiles/prior_stats.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MAS_PRIOR_STATS"] = "1"
import numpy as np
import torch
import bench
from art_tts_b200 import _lib

dev = torch.device("cuda:0")
B, T_X, T_Y, F = 1024, bench.T_X, bench.T_Y, bench.N_FEATS
t_x_np, t_y_np = bench.make_lengths(B, 1000)
t_x, t_y = torch.from_numpy(t_x_np).to(dev), torch.from_numpy(t_y_np).to(dev)
mu_x = torch.randn(B, F, T_X, device=dev)
y = torch.randn(B, F, T_Y, device=dev)
lib = _lib.load()
sbytes = 1024 * 16 * 8
nws = int(lib.mas_workspace_bytes(B, T_X, T_Y)) + sbytes + 64
ws = torch.zeros(nws, dtype=torch.uint8, device=dev)
path = torch.empty(B, T_X, T_Y, device=dev)
dur = torch.empty(B, T_X, dtype=torch.int32, device=dev)
for _ in range(3):
    code = lib.mas_from_prior_f32(_lib.ptr(mu_x), None, _lib.ptr(y), _lib.ptr(t_x), _lib.ptr(t_y),
                                  _lib.ptr(path), 0, _lib.ptr(dur), None, None, None, B, F, T_X, T_Y,
                                  _lib.ptr(ws), nws, 0, _lib.stream_ptr(dev))
    assert code == 0, code
torch.cuda.synchronize()
off = (nws - sbytes) & ~15
t = ws[off:off + sbytes].view(torch.int64).view(1024, 16).cpu().numpy().astype(np.float64)
t = t[t[:, 0] > 0]
print(f"{len(t)} CTAs; utterances/CTA {t[:, 5].mean():.2f}; tiles/CTA {t[:, 6].mean():.1f}")
def row(n, v):
    print(f"{n:44s} {v.mean():12.0f} cyc  {100 * v.mean() / t[:, 0].mean():5.1f}%  per tile {v.mean() / t[:, 6].mean():8.0f}")
row("DP warp total", t[:, 0])
row("  forward (incl. starved of tiles)", t[:, 1])
row("    starved of tiles", t[:, 7])
row("  waiting for a free bit buffer", t[:, 3])
row("backtrack warp: backtrack", t[:, 2])
row("backtrack warp: outputs", t[:, 4])
row("backtrack warp: waiting for forward", t[:, 14])
row("backtrack warp: waiting for zero fill", t[:, 13])
names = ["FMA warp0 total", "  mu_x load + musq (incl. barrier)", "  wait DP (ring stage free)", "  wait y slab", "  compute items"]
for i, n in enumerate(names):
    v = t[:, 8 + i]
    print(f"{n:40s} {v.mean():12.0f} cyc  {100 * v.mean() / t[:, 8].mean():5.1f}%  per tile {v.mean() / t[:, 6].mean():8.0f}")
print("max CTA total", t[:, 0].max(), "min", t[:, 0].min())
