"""One fused call at BASELINE config 4 (B=32, 512 x 4096, F=80): the cluster kernel, for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from art_tts_b200 import monotonic_align
dev = torch.device("cuda", 0)
B, F, T_x, T_y = 32, 80, 512, 4096
g = torch.Generator(device=dev).manual_seed(4)
mu_x = torch.randn(B, F, T_x, device=dev, generator=g)
y = torch.randn(B, F, T_y, device=dev, generator=g)
tx = torch.full((B,), T_x, dtype=torch.int32, device=dev)
ty = torch.full((B,), T_y, dtype=torch.int32, device=dev)
for _ in range(3):
    monotonic_align.maximum_path_from_prior(mu_x, None, y, tx, ty)
torch.cuda.synchronize()
