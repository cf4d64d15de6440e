"""Three drop-in calls at the bench shape (B=1024, 190 x 872, ragged) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from art_tts_b200 import monotonic_align
dev = torch.device("cuda", 0)
tx_np, ty_np = bench.make_lengths(1024, 1000)
tx, ty = torch.from_numpy(tx_np).to(dev), torch.from_numpy(ty_np).to(dev)
value = -(torch.rand(1024, bench.T_X, bench.T_Y, device=dev) * 100 + 50)
for _ in range(3):
    monotonic_align.maximum_path_lengths(value, tx, ty, return_durations=True)
torch.cuda.synchronize()
