// mas_prior.cu -- fused Gaussian log-prior + MAS for sm_100a.
//
// Replaces the block GradTTS/ArtTTS.compute_loss runs under torch.no_grad()
// (src/model/tts.py:483-500 and its copies at :200-214, :776-790, :1067-1081) plus the
// duration sum of tts.py:503-505:
//
//     lp[x,j] = -0.5*sum_f y[f,j]^2 + sum_f mu[f,x]*y[f,j] - 0.5*sum_f mu[f,x]^2 - 0.5*F*log(2*pi)
//     path    = maximum_path(lp, mask);   durations = sum_j path
//
// mas_prior_kernel  PERSISTENT: one CTA per SM walks a list of utterances; three kinds of warps
//                   form a pipeline that keeps running ACROSS utterances:
//                     slab loader  streams 32-frame slabs of y into shared memory (cp.async) and
//                                  clears the dense output path with bulk (TMA) stores;
//                     FMA warps    keep mu_x of the current utterance in shared memory and turn
//                                  each slab into a 32-frame log-prior tile (fp32 FMA on CUDA
//                                  cores: with F <= 80 the contraction is far too thin for tensor
//                                  cores), written into the same swizzled ring the drop-in kernel
//                                  fills from HBM.  Only the band of each tile is computed;
//                     DP warp      consumes the tiles with the single-warp recurrence of
//                                  mas_dp.cuh, backtracks, and writes path/durations.
//                   While the DP warp drains and backtracks utterance k, the other warps are
//                   already loading mu_x and producing tiles of utterance k+1, so the serial
//                   prologue/epilogue of one utterance overlaps the arithmetic of the next.
//                   The T_x x T_y matrix never touches HBM.
// log_prior_kernel  the unfused prior (parity tap `log_prior_out`, and the fallback for shapes
//                   whose operands do not fit in shared memory).
#include <algorithm>
#include <cmath>

#include "mas_dp.cuh"
#include "mas_internal.h"

namespace mas {

// Warp roles.  A warp's scheduler (SMSP) is warp_id % 4.  The frame-sequential DP warp needs
// most of one scheduler's issue slots to run at its natural ~85 cycles/frame, and it cannot be
// given priority over the FMA warps' long independent streams (measured: sharing a scheduler
// with 4 FMA warps serialised the two).  So scheduler 3 is reserved for the latency-bound
// agents -- DP warp (warp 3), slab loader (warp 7), backtrack warp (warp 11) -- plus an optional
// FMA warp (warp 15); schedulers 0-2 run 4 FMA warps each.
constexpr int kFmaPerSmsp = 4;
constexpr int kPriorWarps = 16;
constexpr int kPriorThreads = 32 * kPriorWarps;
constexpr int kDpWarp = 3;
constexpr int kLoaderWarp = 7;
constexpr int kBacktrackWarp = 11;
constexpr int kZeroBytes = 4096;  // zeroed shared buffer behind the bulk stores
constexpr int kSlabs = 4;         // y slabs in flight (own ring, see below)

// FMA-warp index of a warp (or -1): schedulers 0-2 first, then the extras on scheduler 3
__device__ __forceinline__ int fma_warp_index(int warp, int extra, int per_smsp)
{
    const int sm = warp & 3, slot = warp >> 2;
    if (sm != 3) return slot < per_smsp ? slot * 3 + sm : -1;
    return (slot >= 3 && slot - 3 < extra) ? 3 * per_smsp + (slot - 3) : -1;
}

// extra shared memory of the fused kernel (after the ring / bits / bars of FastLayout):
//   mu_s  [F][xrows]        mu_x of the utterance, token axis contiguous (global layout kept)
//   musq  [xrows]           -0.5 * |mu_x|^2 per token
//   yslab [kSlabs][F][32]   ring of 32-frame slabs of y (own ring: a slab is free as soon as
//                           the FMA warps are done with it, not when the DP warp has consumed
//                           the tile made from it)
//   ysq   [kSlabs][32]      -0.5 * |y_j|^2 per frame of the slab (computed by the loader warp)
//   ybar  [kSlabs]          slab-landed mbarriers (cp.async completion -> FMA warps)
//   qbar  [kSlabs]          ysq-ready mbarriers (loader -> FMA warps; only needed by a pass's
//                           epilogue, so the loader's arithmetic is off the critical path)
//   yfree [kSlabs]          slab-consumed mbarriers (FMA warps -> loader); a dedicated pair per
//                           slot keeps every parity wait at most one phase away by construction
//   ctrl  [4] int           progress counters between loader / DP warp / backtrack warp
//   bits2 [nch][xrows]      optional second direction-bit buffer
//   zero  [kZeroBytes]      zeros, source of the bulk stores that clear the output path
struct PriorSmem {
    size_t off_mu, off_musq, off_yslab, off_ysq, off_ybar, off_ctrl, off_zero, off_bits2, total_extra;
};

__host__ __device__ inline PriorSmem prior_smem(int F, int xrows, size_t bits2_bytes = 0)
{
    PriorSmem s;
    s.off_mu = 0;
    s.off_musq = s.off_mu + (size_t)F * xrows * 4;
    s.off_yslab = s.off_musq + (size_t)xrows * 4;
    s.off_ysq = s.off_yslab + (size_t)kSlabs * F * kTileY * 4;
    s.off_ybar = s.off_ysq + (size_t)kSlabs * kTileY * 4;
    s.off_ctrl = s.off_ybar + (size_t)kSlabs * 24;
    s.off_zero = (s.off_ctrl + 16 + 15) & ~(size_t)15;
    s.off_bits2 = s.off_zero + kZeroBytes;
    s.total_extra = s.off_bits2 + bits2_bytes;
    return s;
}

// `second_bits`: room for a second direction-bit buffer (the backtrack warp works on one while the
// DP warp fills the other)
size_t prior_extra_smem(int F, int T_x, int T_y, bool second_bits)
{
    const size_t xrows = (size_t)((T_x + 31) / 32) * 32, nch = (size_t)(T_y + 31) / 32;
    return prior_smem(F, (int)xrows, second_bits ? nch * xrows * 4 : 0).total_extra;
}

// One work item = 32 tokens x 32 frames of the prior by one warp: thread (xg, yg) owns 4 tokens
// x 8 frames, i.e. 32 independent fp32 FMA chains (16 packed FFMA2 per feature) fed by 3
// conflict-free LDS.128.  Accumulation runs over f in ascending order with
// one FMA per term, exactly like log_prior_kernel / lp_cell: all three are bit-identical.
__device__ __forceinline__ void prior_pass(const float *__restrict__ mu_s, const float *__restrict__ musq,
                                           const float *__restrict__ ys, const float *__restrict__ ysq,
                                           float *__restrict__ tile, int F, int xrows, int xbase,
                                           int xlimit, int lane, float cst, const RowMap rm,
                                           uint64_t *qbar, uint32_t qparity)
{
    // tokens xbase .. xbase+31 (xbase % 4 == 0); rows >= xlimit are neither needed nor stored
    const int xg = lane >> 2, yg = lane & 3;
    const int x0 = xbase + 4 * xg;
    uint32_t ma = smem_u32(mu_s + x0);            // walks down the feature axis of mu_s
    const uint32_t mstep = 4u * (uint32_t)xrows;
    uint32_t ya = smem_u32(ys + 8 * yg);          // slab row stride is a constant 128 B
    float acc[4][8];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[r][k] = 0.0f;
    // software-pipelined: the operands of feature f+1 are in flight while feature f is consumed;
    // packed FMAs (two frames per instruction) halve the issue slots of the 32 chains
    auto fma_tile = [&](const float4 &m, const float4 &y0, const float4 &y1) {
        const float mr[4] = {m.x, m.y, m.z, m.w};
        const float yk[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 8; k += 2) ffma2(acc[r][k], acc[r][k + 1], mr[r], yk[k], yk[k + 1]);
    };
    float4 m = lds128(ma), y0 = lds128(ya), y1 = lds128(ya + 16);
#pragma unroll 4
    for (int f = 1; f < F; ++f) {
        ma += mstep;
        ya += 4 * kTileY;
        const float4 mn = lds128(ma), y0n = lds128(ya), y1n = lds128(ya + 16);
        fma_tile(m, y0, y1);
        m = mn;
        y0 = y0n;
        y1 = y1n;
    }
    fma_tile(m, y0, y1);
    mbar_wait(qbar, qparity);  // -0.5|y|^2 of this slab (loader warp), long ready by now
    const float4 qa = *reinterpret_cast<const float4 *>(ysq + 8 * yg);
    const float4 qb = *reinterpret_cast<const float4 *>(ysq + 8 * yg + 4);
    const float q[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int x = x0 + r;
        if (x >= xlimit) continue;
        const float msq = musq[x];
        float o[8];
        // tts.py:495: y_square - y_mu_double + mu_square + const  (y_mu_double == -cross)
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = ((q[k] + acc[r][k]) + msq) + cst;
        const int pr = rm.row(x);  // physical row of token x in the staged tile (mas_dp.cuh)
        float *row = tile + (pr << 5);
        *reinterpret_cast<float4 *>(row + (((2 * yg) ^ (pr & 7)) << 2)) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4 *>(row + (((2 * yg + 1) ^ (pr & 7)) << 2)) = make_float4(o[4], o[5], o[6], o[7]);
    }
}

// Optional coarse per-CTA cycle accounting (PriorArgs::stats, enabled by MAS_PRIOR_STATS=1;
// read back by profiles/prior_stats.py).  clock64 at phase boundaries only: no cost when off.
struct Stat {
    long long acc = 0, t0 = 0;
    bool on;
    __device__ __forceinline__ explicit Stat(bool e) : on(e) {}
    __device__ __forceinline__ void begin() { if (on) t0 = clock64(); }
    __device__ __forceinline__ void end() { if (on) acc += clock64() - t0; }
};

__device__ __forceinline__ void fma_bar(int nthreads)
{
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

template <int XPLMAX>
__global__ void __launch_bounds__(kPriorThreads) mas_prior_kernel(const PriorArgs a)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const FastLayout &L = a.lay;
    const int F = a.F, T_x = a.T_x, NS = L.nstages;
    const int64_t T_y = a.T_y;
    const PriorSmem ps = prior_smem(F, L.xrows);
    float *stages = reinterpret_cast<float *>(smem + L.off_stages);
    int *first = reinterpret_cast<int *>(smem + L.off_first);
    int *dur = reinterpret_cast<int *>(smem + L.off_dur);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.off_bars);
    unsigned char *extra = smem + L.off_bars + 128;
    float *mu_s = reinterpret_cast<float *>(extra + ps.off_mu);
    float *musq = reinterpret_cast<float *>(extra + ps.off_musq);
    float *yslab = reinterpret_cast<float *>(extra + ps.off_yslab);
    float *ysq = reinterpret_cast<float *>(extra + ps.off_ysq);
    uint64_t *ybar = reinterpret_cast<uint64_t *>(extra + ps.off_ybar);
    volatile int *zdone = reinterpret_cast<volatile int *>(extra + ps.off_ctrl);
    uint32_t *zbuf = reinterpret_cast<uint32_t *>(extra + ps.off_zero);
    uint32_t *bits2 = reinterpret_cast<uint32_t *>(extra + ps.off_bits2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nfma = 3 * a.fma_per_smsp + a.extra_fma;
    const int cw = fma_warp_index(warp, a.extra_fma, a.fma_per_smsp);
    const float cst = (float)(-0.5 * 1.8378770664093453 * (double)F);  // -0.5*log(2*pi)*F, tts.py:484

    uint32_t *bits_smem = reinterpret_cast<uint32_t *>(smem + L.off_bits);
    TileRing ring;
    ring.stages = stages;
    ring.full = bars;
    ring.empty = bars + NS;
    ring.nstages = NS;
    ring.stage_floats = L.xrows * kTileY;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&ring.full[s], nfma);  // every FMA warp arrives once per tile
            mbar_init(&ring.empty[s], 1);
        }
        for (int s = 0; s < kSlabs; ++s) {
            mbar_init(&ybar[s], 32);             // one cp.async arrival per loader lane
            mbar_init(&ybar[2 * kSlabs + s], 1);  // qbar: loader lane 0 once ysq is stored
            mbar_init(&ybar[kSlabs + s], nfma);  // yfree: every FMA warp arrives once per tile
        }
        zdone[0] = 0;
        zdone[1] = 0;
        zdone[2] = 0;
        mbar_fence_init();
    }
    for (int i = tid; i < kZeroBytes / 4; i += kPriorThreads) zbuf[i] = 0u;
    fence_proxy_async_smem();  // generic-proxy zeros -> visible to the bulk-copy (async) proxy
    __syncthreads();

    // Every role walks the same utterance list u = blockIdx.x, +gridDim.x, ... (the batch is
    // length-bucketed longest-first by the caller, so this deal is an LPT schedule) and derives
    // the same per-utterance geometry; `g` counts tiles over the CTA's lifetime and indexes the
    // rings, whose barrier phases simply keep running from one utterance to the next.
    auto geometry = [&](int u, int &tx, int &ty, int &ntiles, bool &degenerate) {
        tx = min(max(a.t_x[u], 0), T_x);
        ty = min(max(a.t_y[u], 0), a.T_y);
        degenerate = tx > ty && ty >= 1;
        const bool active = tx >= 1 && ty >= 1 && !degenerate;
        ntiles = active ? (ty + kTileY - 1) / kTileY : 0;
    };

    // ctrl words (shared, volatile): [0] zdone   loader     -> backtrack warp: utterances whose output
    //                                            path has been cleared
    //                                [1] fwd_done DP warp    -> backtrack warp: forward passes finished
    //                                [2] bt_done  backtrack  -> DP warp: direction-bit buffers released
    volatile int *fwd_done = zdone + 1, *bt_done = zdone + 2;
    const int bslots = a.bits_slots;  // direction-bit buffers in shared memory (0 = spilled to global)
    auto bits_of = [&](int u, int k) -> uint32_t * {
        if (!L.bits_in_smem) return a.bits_ws + (size_t)u * L.nch * L.xrows;
        return (k & 1) && bslots == 2 ? bits2 : bits_smem;
    };

    if (warp == kDpWarp) {
        // ======================= DP warp: forward recurrence only =======================
        int g = 0, k = 0;
        Stat st_fwd(a.stats != nullptr), st_w(a.stats != nullptr), st_all(a.stats != nullptr);
        long long dp_wait = 0;
        st_all.begin();
        for (int u = blockIdx.x; u < a.B; u += gridDim.x, ++k) {
            int tx, ty, ntiles;
            bool degenerate;
            geometry(u, tx, ty, ntiles, degenerate);
            float score = 0.0f;
            if (ntiles > 0) {
                // the bit buffer of utterance k is free once utterance k-bslots has been backtracked
                st_w.begin();
                if (bslots > 0)
                    while (*bt_done < k - bslots + 1) __nanosleep(32);
                __threadfence_block();
                st_w.end();
                st_fwd.begin();
                score = prior_forward_dispatch<XPLMAX, true>(ring, bits_of(u, k), L.xrows, tx, ty, lane, g,
                                                       a.stats ? &dp_wait : nullptr);
                g += ntiles;
                st_fwd.end();
                if (lane == 0 && a.score) a.score[u] = score;
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) *fwd_done = k + 1;  // direction bits of utterance k are complete
        }
        st_all.end();
        if (a.stats && lane == 0) {
            long long *o = a.stats + (size_t)blockIdx.x * 16;
            o[0] = st_all.acc; o[1] = st_fwd.acc; o[3] = st_w.acc;
            o[5] = k; o[6] = g; o[7] = dp_wait;
        }
    } else if (warp == kBacktrackWarp) {
        // ======================= backtrack warp: path, durations, frame index =======================
        // Runs one utterance behind the DP warp on the second bit buffer, so the frame-sequential
        // recurrence never pauses for the (equally sequential) backtrack.
        int k = 0;
        Stat st_bt(a.stats != nullptr), st_z(a.stats != nullptr), st_out(a.stats != nullptr),
            st_w(a.stats != nullptr);
        for (int u = blockIdx.x; u < a.B; u += gridDim.x, ++k) {
            int tx, ty, ntiles;
            bool degenerate;
            geometry(u, tx, ty, ntiles, degenerate);
            for (int x = lane; x < T_x; x += 32) dur[x] = 0;
            st_w.begin();
            while (*fwd_done <= k) __nanosleep(32);
            __threadfence_block();
            __syncwarp();
            st_w.end();
            st_bt.begin();
            if (ntiles > 0) {
                if (lane == 0) backtrack_bits(bits_of(u, k), L.xrows, tx, ty, first, dur, L.bits_in_smem != 0);
            } else if (degenerate) {
                if (lane == 0) {  // reference semantics for t_x > t_y: raw prior values (mas_dp.cuh)
                    const float *mub = a.mu_x + (int64_t)u * F * T_x;
                    const float *yb = a.y + (int64_t)u * F * T_y;
                    auto val = [&](int x, int y) { return lp_cell(mub, yb, F, T_x, T_y, x, y, cst); };
                    backtrack_degenerate(val, tx, ty, first, dur);
                    if (a.score) a.score[u] = val(tx - 1, ty - 1);
                }
            } else if (lane == 0 && a.score) {
                a.score[u] = 0.0f;
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) *bt_done = k + 1;  // the DP warp may reuse this bit buffer
            st_bt.end();
            // the output path of utterance k must have been cleared before its 1-cells are written
            st_z.begin();
            if (a.path) {
                while (*zdone <= k) __nanosleep(64);
                __threadfence_block();
            }
            st_z.end();
            st_out.begin();
            char *pb = a.path ? static_cast<char *>(a.path) + (int64_t)u * T_x * T_y * a.path_esize : nullptr;
            write_path_ones(pb, a.durations ? a.durations + (int64_t)u * T_x : nullptr, first, dur, T_x,
                            T_y, a.path_esize, a.one, lane, 32);
            write_frame_idx(a.frame_idx ? a.frame_idx + (int64_t)u * T_y : nullptr, first, dur, T_x, ty,
                            a.T_y, lane, 32);
            __syncwarp();
            st_out.end();
        }
        if (a.stats && lane == 0) {
            long long *o = a.stats + (size_t)blockIdx.x * 16;
            o[2] = st_bt.acc; o[4] = st_out.acc; o[13] = st_z.acc; o[14] = st_w.acc;
        }
    } else if (warp == kLoaderWarp) {
        // ======================= slab loader =======================
        // y[:, 32t..32t+31] -> shared (cp.async, completion signalled straight to the slab barrier),
        // and the zero-fill of the dense output path (bulk stores drained by the copy engine).
        const bool vec16 = (T_y % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.y) & 15) == 0);
        constexpr int kLag = kSlabs - 2;  // slabs copying while an older one is being finished
        int gfin = 0;                     // next slab to finish
        auto finish_slab = [&](int gg) {
            const int s = gg % kSlabs;
            __syncwarp();                 // every lane's copies of this slab have landed
            const float *src = yslab + (size_t)s * F * kTileY + lane;
            float q = 0.0f;
            int f = 0;
            for (; f + 16 <= F; f += 16) {  // 16 loads in flight, then the ordered FMA chain
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = src[(f + i) * kTileY];
#pragma unroll
                for (int i = 0; i < 16; ++i) q = __fmaf_rn(v[i], v[i], q);
            }
            for (; f < F; ++f) {
                const float v = src[f * kTileY];
                q = __fmaf_rn(v, v, q);
            }
            ysq[s * kTileY + lane] = -0.5f * q;  // tts.py:488-490  y_square
            __syncwarp();
            if (lane == 0) mbar_arrive(&ybar[2 * kSlabs + s]);
        };
        int g = 0, k = 0;
        for (int u = blockIdx.x; u < a.B; u += gridDim.x, ++k) {
            int tx, ty, ntiles;
            bool degenerate;
            geometry(u, tx, ty, ntiles, degenerate);
            const float *yb = a.y + (int64_t)u * F * T_y;
            char *pb = a.path ? static_cast<char *>(a.path) + (int64_t)u * T_x * T_y * a.path_esize : nullptr;
            const int64_t pbytes = a.path ? (int64_t)T_x * T_y * a.path_esize : 0;
            const bool zbulk = bulk_zero_ok(pb, pbytes);
            if (zbulk) {
                zero_fill_bulk_part(pb, pbytes, 0, 1, zbuf, kZeroBytes, lane, 32);
                bulk_commit();
            } else {
                zero_fill_part(pb, pbytes, 0, 1, lane, 32);
            }
            // Software pipeline over the CTA-lifetime tile index: slab copies run kLag tiles ahead
            // of the point where a slab is finished (-0.5|y|^2 per frame, then the ready barrier).
            for (int t = 0; t < ntiles; ++t, ++g) {
                const int s = g % kSlabs;
                if (g >= kSlabs)  // slot's previous slab (tile g-kSlabs) fully used by the FMA warps
                    mbar_wait(&ybar[kSlabs + s], ((g / kSlabs) - 1) & 1);
                float *dst = yslab + (size_t)s * F * kTileY;
                const int y0 = t * kTileY;
                if (vec16) {
                    const int c = lane & 7, r = lane >> 3;
                    const int left = ty - (y0 + 4 * c);
                    const uint32_t bytes = left >= 4 ? 16u : (left > 0 ? 4u * left : 0u);
                    const int yo = bytes ? y0 + 4 * c : 0;
                    for (int f = r; f < F; f += 4)
                        cp_async16(dst + f * kTileY + 4 * c, yb + (int64_t)f * T_y + yo, bytes);
                } else {
                    const int y = y0 + lane;
                    const uint32_t bytes = y < ty ? 4u : 0u;
                    const float *src = yb + (y < ty ? y : 0);
                    for (int f = 0; f < F; ++f)
                        cp_async4(dst + f * kTileY + lane, src + (int64_t)f * T_y, bytes);
                }
                cp_async_arrive(&ybar[s]);  // slab-landed barrier fires when this lane's copies arrive
                asm volatile("cp.async.commit_group;" ::: "memory");
                if (g - gfin >= kLag) {  // oldest unfinished slab has had kLag tiles to land
                    asm volatile("cp.async.wait_group %0;" ::"n"(kLag) : "memory");
                    finish_slab(gfin++);
                }
            }
            // drain: finish the slabs still in flight (the FMA warps of THIS utterance need them)
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            while (gfin < g) finish_slab(gfin++);
            // zeros of utterance k are in global memory -> the DP warp may write its 1-cells
            if (zbulk) bulk_wait_all();
            __threadfence_block();
            __syncwarp();
            if (lane == 0) *zdone = k + 1;
        }
    } else if (cw >= 0) {
        // ======================= FMA warps =======================
        // Work items (tile, 32-token pass) are dealt round-robin over the FMA warps,
        // continuing across tiles and utterances.  Every FMA warp waits for every slab and arrives
        // on every tile's `full` barrier (with or without work in it): parity waits are only sound
        // if no waiter can fall two phases behind or run a phase ahead of a barrier.
        const int nft = nfma * 32;
        const int ftid = cw * 32 + lane;
        int g = 0;
        int item0 = 0;  // running item counter: keeps the deal balanced across tiles/utterances
        const bool son = a.stats != nullptr && cw == 0;
        Stat st_mu(son), st_e(son), st_y(son), st_c(son), st_all(son);
        st_all.begin();
        for (int u = blockIdx.x; u < a.B; u += gridDim.x) {
            int tx, ty, ntiles;
            bool degenerate;
            geometry(u, tx, ty, ntiles, degenerate);
            if (ntiles == 0) continue;
            const int npass = (tx + 31) >> 5;
            const int xr = npass * 32;  // token rows any pass may touch
            const float *mub = a.mu_x + (int64_t)u * F * T_x;
            const RowMap rm(tx);
            st_mu.begin();
            fma_bar(nft);  // every FMA warp is done with the previous utterance's mu_s
            for (int i = ftid; i < F * xr; i += nft) {  // all requests in flight at once
                const int f = i / xr, x = i - f * xr;
                cp_async4(mu_s + f * L.xrows + x, mub + (int64_t)f * T_x + (x < tx ? x : 0),
                          x < tx ? 4u : 0u);
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            fma_bar(nft);
            for (int x = ftid; x < xr; x += nft) {
                float s = 0.0f;
                int f = 0;
                for (; f + 16 <= F; f += 16) {
                    float m[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) m[q] = mu_s[(f + q) * L.xrows + x];
#pragma unroll
                    for (int q = 0; q < 16; ++q) s = __fmaf_rn(m[q], m[q], s);
                }
                for (; f < F; ++f) {
                    const float m = mu_s[f * L.xrows + x];
                    s = __fmaf_rn(m, m, s);
                }
                musq[x] = -0.5f * s;  // tts.py:494  mu_square = sum(factor * mu^2)
            }
            fma_bar(nft);
            st_mu.end();
            for (int t = 0; t < ntiles; ++t, ++g) {
                const int s = g % NS, ys = g % kSlabs;
                const int lo = max(0, tx + t * kTileY - ty);
                const int hi = min(tx - 1, t * kTileY + kTileY - 1);
                const int lo4 = lo & ~3;                   // passes start at the band's lower edge
                const int nit = (hi - lo4 + 32) >> 5;      // 32-token passes needed by this tile
                st_e.begin();
                if (g >= NS) mbar_wait(&ring.empty[s], ((g / NS) - 1) & 1);  // DP consumed tile g-NS
                st_e.end();
                st_y.begin();
                mbar_wait(&ybar[ys], (g / kSlabs) & 1);                       // slab has landed
                st_y.end();
                st_c.begin();
                int it = (cw - item0) % nfma;
                if (it < 0) it += nfma;
                for (; it < nit; it += nfma)
                    prior_pass(mu_s, musq, yslab + (size_t)ys * F * kTileY, ysq + ys * kTileY,
                               stages + (size_t)s * ring.stage_floats, F, L.xrows, lo4 + 32 * it, hi + 1,
                               lane, cst, rm, &ybar[2 * kSlabs + ys], (g / kSlabs) & 1);
                item0 = (item0 + nit) % nfma;
                __syncwarp();
                st_c.end();
                if (lane == 0) {
                    mbar_arrive(&ring.full[s]);           // tile g produced (this warp's share)
                    mbar_arrive(&ybar[kSlabs + ys]);      // slab g no longer needed by this warp
                }
            }
        }
        st_all.end();
        if (son && lane == 0) {
            long long *o = a.stats + (size_t)blockIdx.x * 16 + 8;
            o[0] = st_all.acc; o[1] = st_mu.acc; o[2] = st_e.acc; o[3] = st_y.acc; o[4] = st_c.acc;
        }
    }
}

// ------------------------------------------------------------------------------------
// unfused prior: lp[b,x,y] for every cell, same arithmetic order as the fused producers
// (ascending-f FMA chains, ((ysq + cross) + musq) + const) => bit-identical values.
// Block = 4 warps = 128 tokens; it keeps its mu tile in shared memory and walks kLpSlabs
// 32-frame slabs of y, each warp computing 32 tokens x 32 frames per slab with the 4x8
// register tile of prior_pass.
// ------------------------------------------------------------------------------------
constexpr int kLpTokens = 128, kLpSlabs = 8;

__global__ void __launch_bounds__(128) log_prior_kernel(const float *__restrict__ mu_x,
                                                        const float *__restrict__ y, float *lp,
                                                        int F, int T_x, int T_y)
{
    extern __shared__ __align__(16) unsigned char smem[];
    float *mu_t = reinterpret_cast<float *>(smem);          // [F][128]
    float *ys = mu_t + (size_t)F * kLpTokens;               // [F][32]
    float *musq = ys + (size_t)F * kTileY;                  // [128]
    float *ysq = musq + kLpTokens;                          // [32]
    const int b = blockIdx.z, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int xb = blockIdx.y * kLpTokens;
    const float *mub = mu_x + (int64_t)b * F * T_x;
    const float *yb = y + (int64_t)b * F * T_y;
    const float cst = (float)(-0.5 * 1.8378770664093453 * (double)F);
    for (int i = tid; i < F * kLpTokens; i += 128) {
        const int f = i >> 7, x = xb + (i & 127);
        mu_t[i] = (x < T_x) ? __ldg(mub + (int64_t)f * T_x + x) : 0.0f;
    }
    __syncthreads();
    {
        float s = 0.0f;
        for (int f = 0; f < F; ++f) {
            const float m = mu_t[f * kLpTokens + tid];
            s = __fmaf_rn(m, m, s);
        }
        musq[tid] = -0.5f * s;
    }
    const int xg = lane >> 2, yg = lane & 3;
    const int xl = 32 * warp + 4 * xg;   // this thread's first token inside the block
    for (int sl = 0; sl < kLpSlabs; ++sl) {
        const int y0 = (blockIdx.x * kLpSlabs + sl) * kTileY;
        if (y0 >= T_y) break;
        __syncthreads();                 // previous slab fully consumed
        for (int i = tid; i < F * kTileY; i += 128) {
            const int f = i >> 5, yy = y0 + (i & 31);
            ys[i] = (yy < T_y) ? __ldg(yb + (int64_t)f * T_y + yy) : 0.0f;
        }
        __syncthreads();
        if (tid < kTileY) {
            float q = 0.0f;
            for (int f = 0; f < F; ++f) {
                const float v = ys[f * kTileY + tid];
                q = __fmaf_rn(v, v, q);
            }
            ysq[tid] = -0.5f * q;
        }
        __syncthreads();
        uint32_t ma = smem_u32(mu_t + xl);
        uint32_t ya = smem_u32(ys + 8 * yg);
        float acc[4][8];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[r][k] = 0.0f;
#pragma unroll 4
        for (int f = 0; f < F; ++f) {
            const float4 m = lds128(ma);
            const float4 v0 = lds128(ya);
            const float4 v1 = lds128(ya + 16);
            ma += 4 * kLpTokens;
            ya += 4 * kTileY;
            const float mr[4] = {m.x, m.y, m.z, m.w};
            const float yk[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < 8; k += 2) ffma2(acc[r][k], acc[r][k + 1], mr[r], yk[k], yk[k + 1]);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int x = xb + xl + r;
            if (x >= T_x) continue;
            const float msq = musq[xl + r];
            float *dst = lp + ((int64_t)b * T_x + x) * T_y + y0 + 8 * yg;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (y0 + 8 * yg + k < T_y) dst[k] = ((ysq[8 * yg + k] + acc[r][k]) + msq) + cst;
        }
    }
}

int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

cudaError_t launch_from_prior(const PriorArgs &a, cudaStream_t st)
{
    const int xplmax = (a.T_x + 31) / 32;
    void (*k)(const PriorArgs) = nullptr;
    if (xplmax <= 2) k = mas_prior_kernel<2>;
    else if (xplmax <= 4) k = mas_prior_kernel<4>;
    else if (xplmax <= 6) k = mas_prior_kernel<6>;
    else if (xplmax <= 8) k = mas_prior_kernel<8>;
    else if (xplmax <= 12) k = mas_prior_kernel<12>;
    else k = mas_prior_kernel<16>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)a.lay.total);
    if (e != cudaSuccess) return e;
    // persistent: CTAs per SM that fit in shared memory (1 at F=80/T_x=190, 2 at the F=16 config)
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kPriorThreads, a.lay.total) != cudaSuccess ||
        per_sm < 1)
        per_sm = 1;
    const int grid = std::min(a.B, std::max(1, sm_count() - sm_reserve()) * per_sm);
    k<<<grid, kPriorThreads, a.lay.total, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_log_prior(const float *mu_x, const float *y, float *lp, int B, int F, int T_x,
                             int T_y, cudaStream_t st)
{
    const size_t smem = ((size_t)F * (kLpTokens + kTileY) + kLpTokens + kTileY) * sizeof(float);
    if (smem > (size_t)kSmemBudget) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(log_prior_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid((T_y + kTileY * kLpSlabs - 1) / (kTileY * kLpSlabs), (T_x + kLpTokens - 1) / kLpTokens, B);
    log_prior_kernel<<<grid, 128, smem, st>>>(mu_x, y, lp, F, T_x, T_y);
    count_launch();
    return cudaGetLastError();
}

}  // namespace mas
