"""Per-CTA phase timeline of mas_prior_kernel (needs the -DMAS_TIMING build).

  python profiles/prior_timeline.py build     # on the CPU box: art_tts_b200/lib/libmas_sm100_timing.so
  MAS_LIB_PATH=art_tts_b200/lib/libmas_sm100_timing.so python profiles/prior_timeline.py run
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TLIB = os.path.join(ROOT, "art_tts_b200", "lib", "libmas_sm100_timing.so")

if sys.argv[1:] == ["build"]:
    from art_tts_b200 import build
    print(build.build(extra=["-DMAS_TIMING"], out=TLIB))
    sys.exit(0)

import numpy as np
import torch
import bench
from art_tts_b200 import _lib

assert os.environ.get("MAS_LIB_PATH"), "set MAS_LIB_PATH to the timing build"
dev = torch.device("cuda:0")
B, T_X, T_Y, F = 1024, bench.T_X, bench.T_Y, bench.N_FEATS
t_x_np, t_y_np = bench.make_lengths(B, 1000)
t_x, t_y = torch.from_numpy(t_x_np).to(dev), torch.from_numpy(t_y_np).to(dev)
mu_x = torch.randn(B, F, T_X, device=dev)
y = torch.randn(B, F, T_Y, device=dev)
lib = _lib.load()
nws = int(lib.mas_workspace_bytes(B, T_X, T_Y)) + B * 32 * 8 + 64
ws = torch.zeros(nws, dtype=torch.uint8, device=dev)
path = torch.empty(B, T_X, T_Y, device=dev)
dur = torch.empty(B, T_X, dtype=torch.int32, device=dev)
for _ in range(3):
    code = lib.mas_from_prior_f32(_lib.ptr(mu_x), None, _lib.ptr(y), _lib.ptr(t_x), _lib.ptr(t_y),
                                  _lib.ptr(path), 0, _lib.ptr(dur), None, None, None, B, F, T_X, T_Y,
                                  _lib.ptr(ws), nws, 0, _lib.stream_ptr(dev))
    assert code == 0, code
torch.cuda.synchronize()
off = (nws - B * 32 * 8) & ~15
t = ws[off:off + B * 32 * 8].view(torch.int64).view(B, 32).cpu().numpy().astype(np.float64)
tot = t[:, 12] - t[:, 0]
names = {"prologue (mu load, musq)": t[:, 1] - t[:, 0],
         "FMA warp0 loop": t[:, 2] - t[:, 1],
         "  waiting empty (DP)": t[:, 3], "  waiting slab": t[:, 4],
         "loader loop": t[:, 5] - t[:, 1],
         "  cp.async wait": t[:, 6], "  waiting full (FMA)": t[:, 7], "  zero fill": t[:, 8],
         "  finish (ysq)": t[:, 9],
         "DP forward end - FMA end": t[:, 10] - t[:, 2],
         "backtrack": t[:, 11] - t[:, 10],
         "epilogue (ones, durations)": t[:, 12] - t[:, 11],
         "total": tot}
ntiles = (t_y_np + 31) // 32
print(f"B={B}; mean t_x={t_x_np.mean():.0f} t_y={t_y_np.mean():.0f} tiles={ntiles.mean():.1f}")
for k, v in names.items():
    print(f"{k:32s} mean {v.mean():10.0f} cyc  ({100 * v.mean() / tot.mean():5.1f}%)   per tile {v.mean() / ntiles.mean():8.0f}")
for sel, nm in ((slice(0, 8), "longest 8"), (slice(B - 8, B), "shortest 8")):
    print(nm, "total", tot[sel].mean(), "t_x", t_x_np[sel].mean(), "t_y", t_y_np[sel].mean())
