"""H2D bandwidth of the copy shapes the host-buffer entry uses (pinned -> device, one stream):
flat copies vs cudaMemcpy2DAsync with short rows (mu_x: <= 760 B) and long rows (y: 2-3.5 KB)."""
import ctypes
import torch

rt = ctypes.CDLL("libcudart.so.12")
dev = torch.device("cuda:0")
N = 256 << 20
h = torch.empty(N, dtype=torch.uint8).pin_memory()
d = torch.empty(N, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream


def timed(fn, nbytes, label):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{label:46s} {nbytes / 1e6:8.1f} MB  {ms:7.3f} ms  {nbytes / ms / 1e6:6.1f} GB/s")


def flat(n):
    return lambda: rt.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(h.data_ptr()), ctypes.c_size_t(n), 1,
                                      ctypes.c_void_p(st))


def two_d(pitch, width, rows):
    return lambda: rt.cudaMemcpy2DAsync(ctypes.c_void_p(d.data_ptr()), ctypes.c_size_t(pitch), ctypes.c_void_p(h.data_ptr()),
                                        ctypes.c_size_t(pitch), ctypes.c_size_t(width), ctypes.c_size_t(rows), 1,
                                        ctypes.c_void_p(st))


timed(flat(250 << 20), 250 << 20, "flat 250 MB")
timed(flat(32 << 20), 32 << 20, "flat 32 MB (one chunk)")
for pitch, width, rows, label in [(760, 512, 81920, "2D mu_x rows: 512 of 760 B x 81920"),
                                  (760, 640, 81920, "2D mu_x rows: 640 of 760 B x 81920"),
                                  (3488, 2048, 65536, "2D y rows: 2048 of 3488 B x 65536"),
                                  (3488, 3072, 65536, "2D y rows: 3072 of 3488 B x 65536"),
                                  (3488, 1024, 65536, "2D y rows: 1024 of 3488 B x 65536")]:
    timed(two_d(pitch, width, rows), width * rows, label)
