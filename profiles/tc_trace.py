"""Per-tile event timeline of ONE CTA of the fused tensor-core kernel (build with -DMAS_TC_TRACE, MAS_PRIOR_STATS=1).
   python profiles/tc_trace.py          # builds art_tts_b200/lib/libmas_trace.so, runs the bench workload, prints the timeline
Events per CTA-lifetime tile g: loader copy issued / slab ready, MMA pair issue start / commit issued, epilogue (warp 0)
accumulator seen / ring slot acquired / rows written, DP warp 0 and 1 wait begin / start / end."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MAS_PRIOR_STATS"] = "1"
from art_tts_b200 import build as _b
lib = os.path.join(_b.LIBDIR, "libmas_trace.so")
_b.build(extra=["-DMAS_TC_TRACE"] + os.environ.get("MAS_TRACE_EXTRA", "").split(), out=lib)
os.environ["MAS_LIB_PATH"] = lib
import numpy as np, torch
import bench
from art_tts_b200 import _lib
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T_X, T_Y, F = bench.T_X, bench.T_Y, bench.N_FEATS
tx_np, ty_np = bench.make_lengths(B, 1000)
tx, ty = torch.from_numpy(tx_np).to(dev), torch.from_numpy(ty_np).to(dev)
mu_x = torch.randn(B, F, T_X, device=dev); y = torch.randn(B, F, T_Y, device=dev)
l = _lib.load()
sbytes = 1024 * 32 * 8
nws = int(l.mas_workspace_bytes(B, T_X, T_Y)) + sbytes + 64
ws = torch.zeros(nws, dtype=torch.uint8, device=dev)
path = torch.empty(B, T_X, T_Y, device=dev); dur = torch.empty(B, T_X, dtype=torch.int32, device=dev)
for _ in range(2):
    ws.zero_()
    code = l.mas_from_prior_f32(_lib.ptr(mu_x), None, _lib.ptr(y), _lib.ptr(tx), _lib.ptr(ty), _lib.ptr(path), 0, _lib.ptr(dur),
                                None, None, None, B, F, T_X, T_Y, _lib.ptr(ws), nws, 0, _lib.stream_ptr(dev))
    assert code == 0, code
torch.cuda.synchronize()
off = (nws - sbytes) & ~15
t = ws[off:off + sbytes].view(torch.int64).cpu().numpy()
tr = t[148 * 32:].reshape(-1, 16).astype(np.int64)
n = int((tr[:, 7] > 0).sum())
tr = tr[:n]
t0 = tr[tr > 0].min()
names = ["copy", "slab", "mma0", "mmaC", "eSeen", "eSlot", "eDone", "d0S", "d0E", "d1S", "d1E", "d0W", "d1W"]
print(f"{n} tiles traced (CTA 0); utterance lengths of CTA 0: tiles per utterance =",
      [int((ty_np[u] + 31) // 32) for u in range(0, B, 148)])
print("  g " + " ".join(f"{x:>7s}" for x in names) + "   | d0 busy  d0 wait  period")
prev = None
for g in range(n):
    r = tr[g]
    rel = [(int(r[k] - t0) if r[k] > 0 else -1) for k in range(13)]
    busy = rel[8] - rel[7]; wait = rel[7] - rel[11]
    per = (rel[8] - prev) if prev is not None else 0
    prev = rel[8]
    print(f"{g:3d} " + " ".join(f"{x:7d}" for x in rel) + f"   | {busy:7d} {wait:8d} {per:7d}")
# which hand-off is the last one satisfied before each stage starts (averages over the steady part)
s = slice(8, n - 8)
def avg(a): return float(np.mean(a[s]))
print("\naverages over tiles 8..n-8 (cycles):")
print("  slab ready -> MMA pair issue start      ", avg(tr[:, 2] - tr[:, 1]))
print("  MMA issue start -> commit issued        ", avg(tr[:, 3] - tr[:, 2]))
print("  commit issued -> epilogue sees the tile ", avg(tr[:, 4] - tr[:, 3]))
print("  epilogue sees -> ring slot acquired     ", avg(tr[:, 5] - tr[:, 4]))
print("  slot acquired -> rows written           ", avg(tr[:, 6] - tr[:, 5]))
print("  rows written -> DP warp 0 starts        ", avg(tr[:, 7] - tr[:, 6]))
print("  DP warp 0 busy                          ", avg(tr[:, 8] - tr[:, 7]))
print("  DP warp 0 end -> DP warp 1 start        ", avg(tr[:, 9] - tr[:, 8]))
print("  DP warp 1 busy                          ", avg(tr[:, 10] - tr[:, 9]))
print("  copy issued -> slab ready               ", avg(tr[:, 1] - tr[:, 0]))
print("  period (DP warp 0 end to end)           ", avg(np.diff(tr[:, 8], prepend=tr[0, 8])))
