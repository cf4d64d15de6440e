import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from art_tts_b200 import monotonic_align, _lib
dev = torch.device("cuda", 0)
B, T_x, T_y = 64, 190, 872
value = -(torch.rand(B, T_x, T_y, device=dev) * 100 + 50)
tx = torch.full((B,), T_x, dtype=torch.int32, device=dev)
ty = torch.full((B,), T_y, dtype=torch.int32, device=dev)
for _ in range(3):
    monotonic_align.maximum_path_lengths(value, tx, ty, return_durations=True)
torch.cuda.synchronize()
