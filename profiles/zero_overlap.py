"""Does clearing the dense path with the copy engine (cudaMemsetAsync on a side stream, under the previous
step's kernel) beat the kernel's own zero fill?  Bench workload (B=1024, 190 x 872, F=80)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from art_tts_b200 import monotonic_align, _lib
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T_X, T_Y, F = bench.T_X, bench.T_Y, bench.N_FEATS
tx_np, ty_np = bench.make_lengths(B, 1000)
tx, ty = torch.from_numpy(tx_np).to(dev), torch.from_numpy(ty_np).to(dev)
g = torch.Generator(device=dev).manual_seed(1)
sets = [(torch.randn(B, F, T_X, device=dev, generator=g), torch.randn(B, F, T_Y, device=dev, generator=g)) for _ in range(2)]
paths = [torch.zeros(B, T_X, T_Y, device=dev) for _ in range(3)]
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaMemsetAsync.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]
side = torch.cuda.Stream(device=dev)
main = torch.cuda.current_stream()
nbytes = paths[0].numel() * 4


def memset(p, stream, parts=1):
    per = (nbytes // parts + 255) & ~255
    for k in range(parts):
        lo = k * per
        n = min(per, nbytes - lo)
        if n > 0:
            assert rt.cudaMemsetAsync(p.data_ptr() + lo, 0, n, stream.cuda_stream) == 0


def run(mode, iters=30, parts=1):
    for p in paths:
        p.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def step(i):
        mu, y = sets[i & 1]
        p = paths[i % 3]
        if mode == "kernel":          # the kernel clears the path itself (shipping behaviour)
            monotonic_align.maximum_path_from_prior(mu, None, y, tx, ty, path_out=p)
        elif mode == "nozero":        # lower bound: nobody clears (results are wrong after the first round)
            monotonic_align.maximum_path_from_prior(mu, None, y, tx, ty, path_out=p, flags=_lib.FLAG_PATH_ZEROED)
        elif mode == "serial":        # memset in line, then the kernel
            memset(p, main, parts)
            monotonic_align.maximum_path_from_prior(mu, None, y, tx, ty, path_out=p, flags=_lib.FLAG_PATH_ZEROED)
        elif mode in ("overlap", "overlap_fill"):
            # path i was cleared on the side stream while kernel i-1 ran; clear path i+1 under kernel i
            ev = torch.cuda.Event(); ev.record(main)
            side.wait_event(ev)                      # kernel i-2 (last user of path i+1) is done
            if mode == "overlap":
                memset(paths[(i + 1) % 3], side, parts)
            else:
                with torch.cuda.stream(side):
                    paths[(i + 1) % 3].zero_()
            evz = torch.cuda.Event(); evz.record(side)
            monotonic_align.maximum_path_from_prior(mu, None, y, tx, ty, path_out=p, flags=_lib.FLAG_PATH_ZEROED)
            main.wait_event(evz)                     # next kernel needs path i+1 cleared
    for i in range(6):
        step(i)
    torch.cuda.synchronize()
    e0.record()
    for i in range(6, 6 + iters):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def check():
    """overlap mode produces the same path as the shipping kernel"""
    mu, y = sets[0]
    ref, dref = monotonic_align.maximum_path_from_prior(mu, None, y, tx, ty)
    run("overlap", iters=6)
    torch.cuda.synchronize()
    # after run(): step 11 used sets[1], paths[2]; redo one step on a cleared buffer
    paths[0].zero_()
    p, d = monotonic_align.maximum_path_from_prior(mu, None, y, tx, ty, path_out=paths[0], flags=_lib.FLAG_PATH_ZEROED)
    torch.cuda.synchronize()
    return bool(torch.equal(p, ref) and torch.equal(d, dref))


print("same result with a pre-cleared path:", check())
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); t0.record()
for _ in range(10): memset(paths[0], main)
t1.record(); torch.cuda.synchronize()
print(f"cudaMemsetAsync alone: {t0.elapsed_time(t1) / 10:.4f} ms for {nbytes / 1e6:.0f} MB = {nbytes / 1e9 / (t0.elapsed_time(t1) / 10 * 1e-3):.0f} GB/s")
t0.record()
for _ in range(10): paths[0].zero_()
t1.record(); torch.cuda.synchronize()
print(f"torch zero_() alone:   {t0.elapsed_time(t1) / 10:.4f} ms")
for mode, parts in (("kernel", 1), ("nozero", 1), ("serial", 1), ("overlap", 1), ("overlap", 8), ("overlap_fill", 1)):
    print(f"{mode:13s} parts={parts}: {run(mode, parts=parts):.4f} ms per step")
