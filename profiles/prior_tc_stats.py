"""Coarse per-CTA cycle accounting of the tensor-core fused kernel (MAS_PRIOR_STATS=1).
   python profiles/prior_tc_stats.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MAS_PRIOR_STATS"] = "1"
if "--dp" in sys.argv:   # instrumented build: waits inside the DP warps of mas_prior_tc2.cu
    sys.argv.remove("--dp")
    from art_tts_b200 import build as _b
    _exp = os.path.join(_b.LIBDIR, "libmas_exp.so")
    if not os.path.exists(_exp) or os.environ.get("MAS_REBUILD_EXP"):
        _b.build(extra=["-DMAS_TC2_DPSTATS"], out=_exp)
    os.environ["MAS_LIB_PATH"] = _exp
import numpy as np
import torch
import bench
from art_tts_b200 import _lib

dev = torch.device("cuda:0")
CFG4 = "--cfg4" in sys.argv      # BASELINE config 4: B=32, 512 x 4096, full lengths (cluster kernel)
if CFG4:
    sys.argv.remove("--cfg4")
B = int(sys.argv[1]) if len(sys.argv) > 1 else (32 if CFG4 else 1024)
T_X, T_Y, F = (512, 4096, 80) if CFG4 else (bench.T_X, bench.T_Y, bench.N_FEATS)
if CFG4:
    t_x_np, t_y_np = np.full(B, T_X, np.int32), np.full(B, T_Y, np.int32)
else:
    t_x_np, t_y_np = bench.make_lengths(B, 1000)
FLAGS = int(os.environ.get("MAS_STATS_FLAGS", "0"))
t_x, t_y = torch.from_numpy(t_x_np).to(dev), torch.from_numpy(t_y_np).to(dev)
mu_x = torch.randn(B, F, T_X, device=dev)
y = torch.randn(B, F, T_Y, device=dev)
lib = _lib.load()
sbytes = 1024 * 32 * 8
nws = int(lib.mas_workspace_bytes(B, T_X, T_Y)) + sbytes + 64
ws = torch.zeros(nws, dtype=torch.uint8, device=dev)
path = None if os.environ.get("MAS_STATS_NOPATH") else torch.empty(B, T_X, T_Y, device=dev)
dur = torch.empty(B, T_X, dtype=torch.int32, device=dev)
for _ in range(3):
    code = lib.mas_from_prior_f32(_lib.ptr(mu_x), None, _lib.ptr(y), _lib.ptr(t_x), _lib.ptr(t_y),
                                  _lib.ptr(path) if path is not None else None, 0, _lib.ptr(dur), None, None, None, B, F, T_X, T_Y,
                                  _lib.ptr(ws), nws, FLAGS, _lib.stream_ptr(dev))
    assert code == 0, code
torch.cuda.synchronize()
off = (nws - sbytes) & ~15
t = ws[off:off + sbytes].view(torch.int64).view(1024, 32).cpu().numpy().astype(np.float64)
t = t[t[:, 0] > 0]
tiles = t[:, 2].mean()
print(f"{len(t)} CTAs; utterances/CTA {t[:, 3].mean():.2f}; tiles/CTA {tiles:.1f}")
names = {0: "DP warp 0 total", 1: "  starved of tiles", 28: "  waiting for a free bit buffer (backtrack)",
         29: "backtrack warp: backtrack", 30: "backtrack warp: waiting for forward", 31: "backtrack warp: outputs (+ wait zero fill)", 26: "DP warp 1 total", 27: "  starved (tiles / warp 0)",
         4: "loader total", 5: "  wait slab free (MMA done)", 6: "  finish slab (split + ysq)", 7: "  cp.async wait", 23: "    finish: staged -> registers", 24: "    finish: proxy fence + arrive", 25: "  cp.async issue", 
         8: "MMA lane total", 9: "  wait A ready (mu_x -> TMEM)", 10: "  wait slab full", 11: "  wait D buffer empty", 12: "  issue + commit",
         13: "epilogue warp0 total", 16: "  wait D full", 17: "  wait ring stage empty", 18: "  ld + adds + store",
         19: "mu_x mover warp0 total", 20: "  global loads (issue)", 21: "  wait A free (prev MMAs done)", 22: "  split + tcgen05.st", 14: "  zero fill of the dense path (a quarter)"}
if CFG4:
    names.update({1: "  starved of tiles", 14: "  waiting for the boundary (left CTA)", 27: "  starved of tiles", 15: "  waiting for warp 0 / room in the right CTA's ring",
                  28: "backtrack warp: waiting for the right CTA's hand-over", 29: "backtrack warp: windowed walk", 30: "backtrack warp: waiting for forward (incl. zero-fill issue)", 31: "backtrack warp: outputs (+ wait zero fill)"})
if os.environ.get("MAS_LIB_PATH"):
    names.update({13: "DP warp 0: first tcgen05.ld of a tile", 14: "DP warp 0: edge wait (tc2) / zero fill issue", 15: "DP warp 0: later tcgen05.ld waits (tc2) / zero fill wait",
                  16: "DP warp 3: first tcgen05.ld of a tile", 17: "DP warp 3: edge wait", 18: "DP warp 3: later tcgen05.ld waits"})
print(f"slowest CTA: DP warp 0 total {t[:, 0].max():.0f} cyc (mean {t[:, 0].mean():.0f})")
for i, n in names.items():
    v = t[:, i]
    print(f"{n:36s} {v.mean():12.0f} cyc  per tile {v.mean() / tiles:8.0f}")
if os.environ.get("MAS_STATS_PER_CTA"):
    tot = t[:, 0]
    print("per-CTA DP warp 0 total, percentiles 0/25/50/75/100:", np.percentile(tot, [0, 25, 50, 75, 100]).astype(int))
    print("utterances per CTA:", np.bincount(t[:, 3].astype(int)))
    tl = t[:, 2]
    print("tiles per CTA, percentiles 0/25/50/75/100:", np.percentile(tl, [0, 25, 50, 75, 100]).astype(int))
    print("cycles per tile by CTA, percentiles:", np.percentile(tot / tl, [0, 25, 50, 75, 100]).astype(int))
    print("correlation(total, tiles) =", round(float(np.corrcoef(tot, tl)[0, 1]), 3))
    order = np.argsort(tot)
    print("slowest CTAs (index in launch order, tiles, cycles):", [(int(i), int(tl[i]), int(tot[i])) for i in order[-6:]])
    print("fastest CTAs:", [(int(i), int(tl[i]), int(tot[i])) for i in order[:6]])
