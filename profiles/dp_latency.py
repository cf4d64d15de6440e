"""Micro-benchmark: one utterance per SM (B=148), so the kernel time is the serial chain of one
DP warp (forward recurrence + backtrack), not HBM.  Prints cycles per frame at 1965 MHz."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from art_tts_b200 import monotonic_align

dev = torch.device("cuda:0")
for T_x, T_y in [(32, 870), (64, 870), (128, 870), (190, 870), (256, 870), (512, 870), (190, 4096)]:
    B = 148
    value = -(torch.rand(B, T_x, T_y, device=dev) * 100 + 50)
    t_x = torch.full((B,), T_x, dtype=torch.int32, device=dev)
    t_y = torch.full((B,), T_y, dtype=torch.int32, device=dev)
    for want_path in (True, False):
        f = lambda: monotonic_align.maximum_path_lengths(value, t_x, t_y, return_durations=True,
                                                         want_path=want_path)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            f()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"T_x={T_x:4d} T_y={T_y:5d} path={want_path}: {ms*1e3:8.1f} us  "
              f"{ms*1e-3*1.965e9/T_y:7.1f} cycles/frame")
