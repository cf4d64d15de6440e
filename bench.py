#!/usr/bin/env python
"""bench.py -- MAS alignment cells/s on B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--op fused|dropin]

A "step" is one pass of the hot path over one synthetic batch of BASELINE config 5:
B=1024 LJSpeech-shape utterances per GPU (T_text<=190, T_mel<=870 padded to 872, n_feats=80, ragged,
length-bucketed), i.e.  (mu_x, y, lengths) -> log-prior -> MAS -> (path, durations), followed
at N>1 by the NCCL all-gather of the int32 durations.  cells = B*T_text*T_mel (padded).

  value  : whole-job cells/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e    : same op through the public API with pinned HOST buffers; H2D of (mu_x, y, lengths)
           and D2H of (durations, score) inside the timed region, every step
  roofline / cpu_baseline / clocks / gpu_launches : see the task contract and DESIGN.md
  drop_in: the bit-exact maximum_path(value, mask) kernel on the same batch shape
           (value [B,T_x,T_y] fp32 resident; 8 B/cell HBM roofline)

--impl reference times the reference's own host implementation (tts.py:483-505 restated with
torch CPU ops + the reference's compiled Cython kernel from oracle/_ref, OpenMP build, all host
threads) on a bounded sample of the same workload.  That leg, and cpu_baseline, are the only
places this file touches oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "mas_alignment_cells_per_s"
UNIT = "cells/s"
# T_mel: 870 frames padded to 872 by fix_len_compatibility (src/model/utils.py:13-17), which is
# what the reference's collate hands to compute_loss (data_textmel.py:132-153); lengths stay <= 870
B_PER_GPU, T_X, T_Y, T_Y_MAXLEN, N_FEATS = 1024, 190, 872, 870, 80


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--op", default="fused", choices=["fused", "dropin"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU)
    ap.add_argument("--cpu-sample", type=int, default=64, help="utterances in the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-dropin", action="store_true")
    ap.add_argument("--e2e-chunk", type=int, default=0, help="utterances per H2D chunk (0 = library default)")
    ap.add_argument("--e2e-no-trim", action="store_true", help="copy whole padded rows")
    # default "serial": measured on 2 GPUs, the overlapped gather is 3 % faster when it works (0.347 vs
    # 0.357 ms) but the fused kernel is persistent with one CTA per SM, and whenever the collective's CTAs
    # still hold an SM at the next launch the CTAs that do not fit run as a second wave (0.78-1.6 ms observed)
    ap.add_argument("--gather", default="auto", choices=["auto", "overlap", "serial", "p2p"],
                    help="N>1: all-gather of step i in line after its kernel, on a side stream under the kernel of step "
                         "i+1, or done by the fused kernel itself over NVLink peer memory (+ a barrier); auto = p2p when "
                         "every rank can set it up (fused op, tensor-core engine, symmetric memory), else serial")
    ap.add_argument("--reserve-sms", type=int, default=-1,
                    help="SMs the persistent kernel leaves to the overlapped collective (default 0: measured "
                         "at N=8, reserving 8 SMs costs 8%% and the gather overlaps anyway)")
    ap.add_argument("--engine", default="auto", choices=["auto", "tensor", "cuda"],
                    help="prior engine of the fused kernel (auto: 3xTF32 tensor cores for F >= 32)")
    return ap.parse_args()


def make_lengths(B, seed):
    """SURVEY.md 8(d) config 1/5 recipe: t_x ~ U{60..190}, t_y = min(870, 4*t_x + U{0..99}),
    then length-bucketed (sorted by work, longest first)."""
    rng = np.random.default_rng(seed)
    t_x = rng.integers(60, T_X + 1, B).astype(np.int32)
    t_y = np.minimum(T_Y_MAXLEN, 4 * t_x + rng.integers(0, 100, B)).astype(np.int32)
    t_x[0], t_y[0] = T_X, T_Y_MAXLEN
    order = np.argsort(-(t_x.astype(np.int64) * t_y), kind="stable")
    return t_x[order], t_y[order]


def config_dict(args, world):
    return {
        "workload": "config5: B=1024/GPU LJSpeech-shape length-bucketed, fused prior+MAS+durations"
                    if args.op == "fused" else
                    "config5: B=1024/GPU LJSpeech-shape length-bucketed, maximum_path(value, mask)",
        "op": args.op, "engine": args.engine, "batch_per_gpu": args.batch, "global_batch": args.batch * world,
        "T_text": T_X, "T_mel": T_Y, "n_feats": N_FEATS, "ragged": True,
        "cells_definition": "B*T_text*T_mel (padded)",
        "l2_policy": "inputs+outputs per step (>= 1 GB) exceed the 126 MB L2; no explicit flush",
        "collective": ("all_gather(durations int32 [B,T_text]), " +
                       ("auto: fused over NVLink peer memory, else NCCL in line" if args.gather == "auto" else args.gather)
                       if world > 1 else "none"),
    }


# ------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            p = [x.strip() for x in s.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0]))
                mx = float(p[1])
            except ValueError:
                continue
            for n, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline (the only users of oracle/)
# ------------------------------------------------------------------------------------------
def cpu_mas_baseline(t_x, t_y, n_sample, budget_s=12.0):
    """maximum_path_c (the reference's Cython kernel, OpenMP build) on host arrays."""
    from oracle import mas_oracle, ref
    n = min(n_sample, len(t_x))
    rng = np.random.default_rng(123)
    value = -(rng.random((n, T_X, T_Y), dtype=np.float32) * 100 + 50)
    tx, ty = np.ascontiguousarray(t_x[:n]), np.ascontiguousarray(t_y[:n])
    kind = "reference" if ref.available("omp") else "port"
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    best, t_end, reps = None, time.perf_counter() + budget_s, 0
    while reps < 3 or (time.perf_counter() < t_end and reps < 20):
        v = value.copy()
        p = np.zeros(v.shape, np.int32)
        t0 = time.perf_counter()
        if kind == "reference":
            ref.maximum_path_c(p, v, tx, ty, kind="omp")
        else:
            mas_oracle.maximum_path_c(p, v, tx, ty, n_threads=cores)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        reps += 1
    cells = n * T_X * T_Y
    return {"value": cells / best, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"maximum_path_c only, B={n} of the workload's utterances, "
                      f"{T_X}x{T_Y} padded, best of {reps}"}


def run_reference(args):
    """--impl reference: host implementation of the same op (prior + MAS + durations)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import prior_torch, ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    n = min(args.cpu_sample, args.batch)
    t_x, t_y = make_lengths(args.batch, 1000)
    t_x, t_y = t_x[:: max(1, args.batch // n)][:n], t_y[:: max(1, args.batch // n)][:n]
    g = torch.Generator().manual_seed(0)
    x_mask = (torch.arange(T_X)[None, :] < torch.from_numpy(t_x)[:, None]).float().unsqueeze(1)
    y_mask = (torch.arange(T_Y)[None, :] < torch.from_numpy(t_y)[:, None]).float().unsqueeze(1)
    mu_x = torch.randn(n, N_FEATS, T_X, generator=g) * x_mask
    y = torch.randn(n, N_FEATS, T_Y, generator=g) * y_mask
    kind_used = None

    def step():
        nonlocal kind_used
        if args.op == "fused":
            _, dur, kind_used = prior_torch.prior_mas_block(mu_x, y, x_mask, y_mask)
        else:
            attn_mask = (x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)).squeeze(1)
            _, kind_used = prior_torch.maximum_path_host(value, attn_mask)

    value = prior_torch.log_prior_block(mu_x, y) if args.op == "dropin" else None
    for _ in range(max(1, min(args.warmup, 3))):
        step()
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        step()
        done += 1
        if time.perf_counter() - t0 > 150:   # keep the whole run within a few minutes
            break
    dt = time.perf_counter() - t0
    cells = n * T_X * T_Y
    val = cells * done / dt
    kind = "reference" if kind_used in ("omp", "serial") else "port"
    sample = (f"B={n} utterances of the workload per step (strided sample of the length-bucketed "
              f"batch), torch-CPU log-prior block + maximum_path wrapper + Cython/OpenMP kernel"
              if args.op == "fused" else
              f"B={n} utterances per step, maximum_path wrapper + Cython/OpenMP kernel")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * dt / done,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_dict(args, max(1, args.gpus)),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from art_tts_b200 import _lib, monotonic_align
    from art_tts_b200 import build as mas_build

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if not os.path.exists(mas_build.LIB):
        raise SystemExit("libmas_sm100.so missing: run __graft_entry__.build() first")
    _lib.load()

    B = args.batch
    cells = B * T_X * T_Y
    t_x_np, t_y_np = make_lengths(B, 1000 + rank)
    t_x = torch.from_numpy(t_x_np).to(dev)
    t_y = torch.from_numpy(t_y_np).to(dev)
    valid_cells = int((t_x_np.astype(np.int64) * t_y_np).sum())
    torch.manual_seed(rank)
    xm = (torch.arange(T_X, device=dev)[None, :] < t_x[:, None]).float()
    ym = (torch.arange(T_Y, device=dev)[None, :] < t_y[:, None]).float()
    mu_x = torch.randn(B, N_FEATS, T_X, device=dev) * xm[:, None, :]
    y = torch.randn(B, N_FEATS, T_Y, device=dev) * ym[:, None, :]
    dur_all = torch.empty(world * B, T_X, dtype=torch.int32, device=dev) if world > 1 else None
    gatherer = None
    reserve = 0
    if world > 1 and args.gather == "overlap":
        from art_tts_b200.distributed import DurationGatherer
        gatherer = DurationGatherer(B, T_X, dev)   # gather of step i overlaps the kernel of step i+1
        # the fused kernel is persistent (one CTA per SM): leave a few SMs to the collective's CTAs
        reserve = 0 if args.reserve_sms < 0 else args.reserve_sms
        _lib.check(_lib.load().mas_set_sm_reserve(reserve), "mas_set_sm_reserve")

    peer = None
    if world > 1 and args.gather in ("p2p", "auto"):
        from art_tts_b200.distributed import PeerDurationGather
        can = args.op == "fused" and args.engine != "cuda" and PeerDurationGather.supported(B, N_FEATS, T_X, T_Y)
        if args.gather == "p2p" and not can:
            raise SystemExit("--gather p2p: only the tensor-core engine of the fused op writes peer memory")
        # every rank must take the same branch: agree before the (collective) rendezvous and after it
        flag = torch.tensor([1 if can else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()):
            try:
                peer = PeerDurationGather(B, T_X, dev)
            except Exception as e:   # no symmetric memory on this box: NCCL gather in line
                if args.gather == "p2p":
                    raise
                print(f"[bench] rank {rank}: peer-memory gather unavailable ({type(e).__name__}: {e}); using NCCL",
                      file=sys.stderr)
            flag = torch.tensor([1 if peer is not None else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if not int(flag.item()) and peer is not None:
                peer.close()
                peer = None
        args.gather = "p2p" if peer is not None else ("serial" if args.gather == "auto" else args.gather)

    value = None
    if args.op == "dropin" or not args.no_dropin:
        value = -(torch.rand(B, T_X, T_Y, device=dev) * 100 + 50)

    eng_flags = {"auto": 0, "tensor": _lib.FLAG_FORCE_TENSOR, "cuda": _lib.FLAG_NO_TENSOR}[args.engine]
    tensor_engine = args.engine == "tensor" or (args.engine == "auto" and N_FEATS >= 32)

    def step_fused():
        path, dur = monotonic_align.maximum_path_from_prior(mu_x, None, y, t_x, t_y, flags=eng_flags)
        if peer is not None:
            peer.finish()            # the kernel wrote every rank's buffer; the ranks only meet here
        elif gatherer is not None:
            gatherer.gather(dur)
        elif world > 1:
            dist.all_gather_into_tensor(dur_all, dur)
        return path, dur

    def step_dropin():
        path, dur = monotonic_align.maximum_path_lengths(value, t_x, t_y, return_durations=True)
        if gatherer is not None:
            gatherer.gather(dur)
        elif world > 1:
            dist.all_gather_into_tensor(dur_all, dur)
        return path, dur

    step = step_fused if args.op == "fused" else step_dropin
    if world > 1:   # bring up every NCCL channel/connection before anything is timed
        for _ in range(8):
            dist.all_gather_into_tensor(dur_all, dur_all[:B])
        torch.cuda.synchronize()
        dist.barrier()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if gatherer is not None:
            gatherer.wait()          # the last gathers are inside the timed region
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = _lib.launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, launches

    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    ms_step, launches = timed(step, args.steps, max(3, args.warmup))
    clk = clocks.stop() if rank == 0 else None
    value_cells = cells * world / (ms_step * 1e-3)

    # ---- per-kernel time for the roofline: CUDA events around the kernel launch alone
    def kernel_only():
        if args.op == "fused":
            monotonic_align.maximum_path_from_prior(mu_x, None, y, t_x, t_y, flags=eng_flags)
        else:
            monotonic_align.maximum_path_lengths(value, t_x, t_y, return_durations=True)

    k_ms, _ = timed(kernel_only, args.steps, 2)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    if args.op == "fused":
        # 4*F*(T_x+T_y) in + 4*T_x durations + 4 B/cell dense path out, per utterance
        alg_bytes = B * (4 * N_FEATS * (T_X + T_Y) + 4 * T_X + 4 * T_X * T_Y)
    else:
        alg_bytes = 8 * cells   # read value fp32 once + write dense fp32 path once
    ach = alg_bytes / (k_ms * 1e-3) / 1e9
    kname = ("mas_prior_tc_kernel" if tensor_engine else "mas_prior_kernel") if args.op == "fused" \
        else "mas_fast_kernel"
    # measured DRAM bytes per launch of this kernel on this workload (ncu --set full,
    # dram__bytes_read.sum + dram__bytes_write.sum; profiles/r1_traffic.json names the capture)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            tj = json.load(f)
        if B == B_PER_GPU:
            traffic = tj.get(kname, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "kernel": kname, "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "frac_of_nominal_8tbs": ach / 8000.0}   # also quoted against the nominal ~8 TB/s
    if args.op == "fused" and tensor_engine:
        roofline["engine"] = ("tcgen05.mma kind::tf32, 3xTF32 split (fp32-level accuracy), mu_x in TMEM; "
                              "the CUDA-core engine (--engine cuda) is the fp32 FMA variant")
    if args.op == "fused" and not tensor_engine:
        # the binding resource at F=80 is the fp32 FMA pipe, not HBM: report it too
        sm_mhz = (clk or {}).get("sm_mhz") or 1965.0
        fma_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        flops = 2.0 * N_FEATS * valid_cells
        roofline["fp32_alu"] = {"achieved_tflops": flops / (k_ms * 1e-3) / 1e12,
                                "peak_tflops_at_observed_clock": fma_peak,
                                "frac": flops / (k_ms * 1e-3) / 1e12 / fma_peak,
                                "flops_counted": "2*F per VALID cell (t_x*t_y)"}

    # ---- drop-in kernel on the same shape (secondary figure, resident value tensor)
    drop = None
    if args.op == "fused" and not args.no_dropin:
        d_ms, _ = timed(lambda: monotonic_align.maximum_path_lengths(value, t_x, t_y,
                                                                     return_durations=True),
                        args.steps, 3)
        d_ach = 8 * cells / (d_ms * 1e-3) / 1e9
        drop = {"value": cells / (d_ms * 1e-3), "unit": UNIT, "ms_per_step": d_ms,
                "roofline": {"bound": "hbm", "achieved": d_ach, "peak": hbm_peak, "unit": "GB/s",
                             "frac": d_ach / hbm_peak,
                             "traffic": (tj.get("mas_fast_kernel", {}).get("dram_bytes_per_launch")
                                         if traffic is not None else None),
                             "kernel": "mas_fast_kernel", "algorithmic_bytes_per_launch": 8 * cells}}

    # ---- e2e: pinned host buffers in, durations + score out, every step
    e2e = None
    if not args.no_e2e:
        if args.op == "fused":
            h_in = [mu_x.cpu().pin_memory(), y.cpu().pin_memory(), t_x.cpu().pin_memory(),
                    t_y.cpu().pin_memory()]
        else:
            h_in = [value.cpu().pin_memory(), t_x.cpu().pin_memory(), t_y.cpu().pin_memory()]
        d_in = [torch.empty_like(h, device=dev) for h in h_in] if args.op != "fused" else []
        h_dur = torch.empty(B, T_X, dtype=torch.int32).pin_memory()
        h_score = torch.empty(B, dtype=torch.float32).pin_memory()
        h2d = sum(h.numel() * h.element_size() for h in h_in)
        d2h = h_dur.numel() * 4 + h_score.numel() * 4

        moved = [h2d]

        def e2e_step():
            if args.op == "fused":
                # host-buffer entry point: trimmed, chunked H2D overlapped with the kernels,
                # D2H of durations + score enqueued behind the last chunk
                path, dur, score, moved[0] = monotonic_align.maximum_path_from_prior_host(
                    h_in[0], h_in[1], h_in[2], h_in[3], dev, chunk=args.e2e_chunk,
                    durations_host=h_dur, score_host=h_score,
                    flags=(_lib.FLAG_HOST_NO_TRIM if args.e2e_no_trim else 0) | eng_flags)
                if peer is not None:
                    peer.finish()
                elif world > 1:
                    dist.all_gather_into_tensor(dur_all, dur)
            else:
                for h, d in zip(h_in, d_in):
                    d.copy_(h, non_blocking=True)
                path, dur, score = monotonic_align.maximum_path_lengths(
                    d_in[0], d_in[1], d_in[2], return_durations=True, return_score=True)
                if world > 1:
                    dist.all_gather_into_tensor(dur_all, dur)
                h_dur.copy_(dur, non_blocking=True)
                h_score.copy_(score, non_blocking=True)
            torch.cuda.current_stream().synchronize()   # the caller reads the result every step

        e_ms, _ = timed(e2e_step, max(3, args.steps // 2), 2)
        e2e = {"value": cells * world / (e_ms * 1e-3), "unit": UNIT, "ms_per_step": e_ms,
               "h2d_bytes_per_step": int(moved[0]), "d2h_bytes_per_step": d2h,
               "padded_input_bytes": h2d,
               "api": "art_tts_b200.monotonic_align.maximum_path_from_prior_host" if args.op == "fused"
                      else "art_tts_b200.monotonic_align.maximum_path_lengths"}

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_mas_baseline(t_x_np, t_y_np, args.cpu_sample)

    line = {"metric": METRIC, "value": value_cells, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_dict(args, world),
            "valid_cells_per_s": valid_cells * world / (ms_step * 1e-3),
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clk, "e2e": e2e,
            "gpu_launches": int(launches), "drop_in": drop}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
