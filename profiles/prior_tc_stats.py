"""Coarse per-CTA cycle accounting of the tensor-core fused kernel (MAS_PRIOR_STATS=1).
   python profiles/prior_tc_stats.py"""
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
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T_X, T_Y, F = bench.T_X, bench.T_Y, bench.N_FEATS
t_x_np, t_y_np = bench.make_lengths(B, 1000)
t_x, t_y = torch.from_numpy(t_x_np).to(dev), torch.from_numpy(t_y_np).to(dev)
mu_x = torch.randn(B, F, T_X, device=dev)
y = torch.randn(B, F, T_Y, device=dev)
lib = _lib.load()
sbytes = 1024 * 32 * 8
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
t = ws[off:off + sbytes].view(torch.int64).view(1024, 32).cpu().numpy().astype(np.float64)
t = t[t[:, 0] > 0]
tiles = t[:, 2].mean()
print(f"{len(t)} CTAs; utterances/CTA {t[:, 3].mean():.2f}; tiles/CTA {tiles:.1f}")
names = {0: "DP warp 0 total", 1: "  starved of tiles", 26: "DP warp 1 total", 27: "  starved (tiles / warp 0)",
         4: "loader total", 5: "  wait slab free (MMA done)", 6: "  finish slab (split + ysq)", 7: "  cp.async wait", 23: "    finish: staged -> registers", 24: "    finish: proxy fence + arrive", 25: "  cp.async issue",
         8: "MMA lane total", 9: "  wait A ready (mu_x -> TMEM)", 10: "  wait slab full", 11: "  wait D buffer empty", 12: "  issue + commit",
         13: "epilogue warp0 total", 16: "  wait D full", 17: "  wait ring stage empty", 18: "  ld + adds + store",
         19: "mu_x mover warp0 total", 20: "  global loads (issue)", 21: "  wait A free (prev MMAs done)", 22: "  split + tcgen05.st"}
for i, n in names.items():
    v = t[:, i]
    print(f"{n:36s} {v.mean():12.0f} cyc  per tile {v.mean() / tiles:8.0f}")
