"""Time the consumers of the alignment (csrc/mas_align.cu) at the bench shape and
quote each against its own HBM roofline (algorithmic bytes / measured copy peak).
   python profiles/align_bench.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from art_tts_b200 import alignment, monotonic_align

dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)
B, T_X, T_Y, F, OUT = 1024, bench.T_X, bench.T_Y, bench.N_FEATS, 172      # out_size of params_v2.py:61
t_x_np, t_y_np = bench.make_lengths(B, 1000)
tx, ty = torch.from_numpy(t_x_np).to(dev), torch.from_numpy(t_y_np).to(dev)
mu_x = torch.randn(B, F, T_X, device=dev)
y = torch.randn(B, F, T_Y, device=dev)
logw = torch.randn(B, 1, T_X, device=dev)
dur, fidx = monotonic_align.maximum_path_from_prior(mu_x, None, y, tx, ty, want_path=False, return_frame_idx=True)
offs, lens = alignment.crop_offsets(t_y_np.tolist(), OUT)
off = torch.tensor(offs, dtype=torch.int32, device=dev)
seg = torch.tensor(lens, dtype=torch.int32, device=dev)
y_cut = alignment.crop(y, off, seg, OUT)


def timeit(f, n=20, w=3):
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


mu_g = mu_x.clone().requires_grad_(True)


def fwd_bwd_full():
    mu_g.grad = None
    m, pl = alignment.aligned_mu_y_and_prior_loss(mu_g, y, fidx, None, ty, T_Y)
    (pl + m.sum() * 0).backward()


def fwd_bwd_cut():
    mu_g.grad = None
    m, pl = alignment.aligned_mu_y_and_prior_loss(mu_g, y_cut, fidx, off, seg, OUT)
    pl.backward()


rows = [
    ("frame_index (durations -> idx)", lambda: alignment.frame_index(dur, tx, ty, T_Y), 4 * B * (T_X + T_Y)),
    ("duration loss fwd (+grad unit)", lambda: alignment.duration_loss_from_durations(logw, dur, tx), 4 * B * T_X * 3),
    ("crop y -> [B,F,172]", lambda: alignment.crop(y, off, seg, OUT), 8 * B * F * OUT),
    ("attn_cut from frame idx [B,T_x,172]", lambda: alignment.path_segment(fidx, off, seg, T_X, OUT), 4 * B * T_X * OUT),
    ("mu_y gather + prior loss, crop 172", lambda: alignment.aligned_mu_y_and_prior_loss(mu_x, y_cut, fidx, off, seg, OUT),
     4 * B * F * (2 * OUT) + 4 * B * OUT),
    ("mu_y gather + prior loss, full T_y", lambda: alignment.aligned_mu_y_and_prior_loss(mu_x, y, fidx, None, ty, T_Y),
     4 * B * F * (T_X + 2 * T_Y)),
    ("  fwd + bwd (segmented sum), crop 172", fwd_bwd_cut, 4 * B * F * (3 * OUT + 2 * T_X)),
    ("  fwd + bwd (segmented sum), full T_y", fwd_bwd_full, 4 * B * F * (4 * T_Y + 3 * T_X)),
]
print(f"B={B} T_x={T_X} T_y={T_Y} F={F}; HBM peak {peak:.0f} GB/s (measured copy)")
print(f"{'kernel(s)':42s} {'ms':>8s} {'alg. MB':>9s} {'GB/s':>8s} {'of peak':>8s}")
for name, f, nbytes in rows:
    ms = timeit(f)
    print(f"{name:42s} {ms:8.4f} {nbytes / 1e6:9.1f} {nbytes / ms / 1e6:8.0f} {nbytes / ms / 1e6 / peak:8.2f}")
