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
// agents -- DP warp (warp 3) and slab loader (warp 7) -- plus an optional, tunable number of
// FMA warps (warps 11, 15); schedulers 0-2 run 4 FMA warps each.
constexpr int kFmaPerSmsp = 4;
constexpr int kPriorWarps = 16;
constexpr int kPriorThreads = 32 * kPriorWarps;
constexpr int kDpWarp = 3;
constexpr int kLoaderWarp = 7;
constexpr int kZeroBytes = 4096;  // zeroed shared buffer behind the bulk stores
constexpr int kSlabs = 4;         // y slabs in flight (own ring, see below)

// FMA-warp index of a warp (or -1): schedulers 0-2 first, then the extras on scheduler 3
__device__ __forceinline__ int fma_warp_index(int warp, int extra)
{
    const int sm = warp & 3, slot = warp >> 2;
    if (sm != 3) return slot * 3 + sm;
    return (slot >= 2 && slot - 2 < extra) ? 3 * kFmaPerSmsp + (slot - 2) : -1;
}

// extra shared memory of the fused kernel (after the ring / bits / bars of FastLayout):
//   mu_s  [F][xrows]        mu_x of the utterance, token axis contiguous (global layout kept)
//   musq  [xrows]           -0.5 * |mu_x|^2 per token
//   yslab [kSlabs][F][32]   ring of 32-frame slabs of y (own ring: a slab is free as soon as
//                           the FMA warps are done with it, not when the DP warp has consumed
//                           the tile made from it)
//   ybar  [kSlabs]          slab-ready mbarriers (loader -> FMA warps)
//   yfree [kSlabs]          slab-consumed mbarriers (FMA warps -> loader); a dedicated pair per
//                           slot keeps every parity wait at most one phase away by construction
//   ctrl  [4] int           zero-fill progress counter (loader -> DP warp)
//   zero  [kZeroBytes]      zeros, source of the bulk stores that clear the output path
struct PriorSmem {
    size_t off_mu, off_musq, off_yslab, off_ybar, off_ctrl, off_zero, total_extra;
};

__host__ __device__ inline PriorSmem prior_smem(int F, int xrows)
{
    PriorSmem s;
    s.off_mu = 0;
    s.off_musq = s.off_mu + (size_t)F * xrows * 4;
    s.off_yslab = s.off_musq + (size_t)xrows * 4;
    s.off_ybar = s.off_yslab + (size_t)kSlabs * F * kTileY * 4;
    s.off_ctrl = s.off_ybar + (size_t)kSlabs * 16;
    s.off_zero = (s.off_ctrl + 16 + 15) & ~(size_t)15;
    s.total_extra = s.off_zero + kZeroBytes;
    return s;
}

size_t prior_extra_smem(int F, int T_x)
{
    return prior_smem(F, ((T_x + 31) / 32) * 32).total_extra;
}

// one cell of the prior, same operation order as the producers / log_prior_kernel
__device__ __forceinline__ float lp_cell(const float *mub, const float *yb, int F, int T_x,
                                         int64_t T_y, int x, int y, float cst)
{
    float ysq = 0.0f, c = 0.0f, msq = 0.0f;
    for (int f = 0; f < F; ++f) {
        const float m = __ldg(mub + (int64_t)f * T_x + x);
        const float v = __ldg(yb + (int64_t)f * T_y + y);
        ysq = __fmaf_rn(v, v, ysq);
        c = __fmaf_rn(m, v, c);
        msq = __fmaf_rn(m, m, msq);
    }
    return ((-0.5f * ysq + c) + -0.5f * msq) + cst;
}

template <int XPLMAX>
__device__ __forceinline__ float prior_forward_dispatch(const TileRing &ring, uint32_t *bits,
                                                        int xrows, int tx, int ty, int lane, int g0)
{
    const int xpl = (tx + 31) >> 5;
#define MAS_CASE(N)                                                                           \
    case N:                                                                                   \
        if constexpr (N <= XPLMAX) return dp_forward<N>(ring, bits, xrows, tx, ty, lane, g0); \
        break;
    switch (xpl) {
        MAS_CASE(1) MAS_CASE(2) MAS_CASE(3) MAS_CASE(4) MAS_CASE(5) MAS_CASE(6) MAS_CASE(7)
        MAS_CASE(8) MAS_CASE(9) MAS_CASE(10) MAS_CASE(11) MAS_CASE(12) MAS_CASE(13) MAS_CASE(14)
        MAS_CASE(15) MAS_CASE(16)
    default: break;
    }
#undef MAS_CASE
    return 0.0f;
}

// One work item = 32 tokens x 16 frames of the prior by one warp: thread (xg, yg) owns 4 tokens
// x 4 frames, i.e. 16 independent fp32 FMA chains fed by 2 LDS.128 per feature (both
// conflict-free).  Items are half the size of a full 32x32 pass so that a tile splits into
// enough pieces to keep every FMA warp busy.  Accumulation runs over f in ascending order with
// one FMA per term, exactly like log_prior_kernel / lp_cell: all three are bit-identical.
__device__ __forceinline__ void prior_item(const float *__restrict__ mu_s, const float *__restrict__ musq,
                                           const float *__restrict__ ys, float *__restrict__ tile, int F,
                                           int xrows, int p, int h, int lane, float cst, const RowMap rm)
{
    const int xg = lane >> 2, yg = lane & 3;
    const int x0 = 32 * p + 4 * xg;
    const int s0 = 16 * h + 4 * yg;             // first of this thread's 4 frames
    const float *mp = mu_s + x0;
    const float *yp = ys + s0;
    // -0.5*|y_j|^2 (tts.py:488-490): of the 8 lanes that share these 4 frames (same yg), lanes
    // xg = 0..3 each accumulate ONE of them alongside the main loop; shuffles distribute them.
    const float *yq = ys + s0 + (xg & 3);
    float qsum = 0.0f;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[r][k] = 0.0f;
#pragma unroll 4
    for (int f = 0; f < F; ++f) {
        const float4 m = *reinterpret_cast<const float4 *>(mp + (size_t)f * xrows);
        const float4 yv = *reinterpret_cast<const float4 *>(yp + f * kTileY);
        const float yo = yq[f * kTileY];
        const float mr[4] = {m.x, m.y, m.z, m.w};
        const float yk[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[r][k] = __fmaf_rn(mr[r], yk[k], acc[r][k]);
        qsum = __fmaf_rn(yo, yo, qsum);
    }
    qsum *= -0.5f;
    float q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] = __shfl_sync(kFull, qsum, (k << 2) | yg);  // lane (xg=k, yg)
    const int chunk = 4 * h + yg;               // 16-byte chunk of the 128-byte tile row
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int x = x0 + r;
        const float msq = musq[x];
        // tts.py:495: y_square - y_mu_double + mu_square + const  (y_mu_double == -cross)
        const float o0 = ((q[0] + acc[r][0]) + msq) + cst, o1 = ((q[1] + acc[r][1]) + msq) + cst;
        const float o2 = ((q[2] + acc[r][2]) + msq) + cst, o3 = ((q[3] + acc[r][3]) + msq) + cst;
        const int pr = rm.row(x);  // physical row of token x in the staged tile (mas_dp.cuh)
        *reinterpret_cast<float4 *>(tile + (pr << 5) + ((chunk ^ (pr & 7)) << 2)) =
            make_float4(o0, o1, o2, o3);
    }
}

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
    uint64_t *ybar = reinterpret_cast<uint64_t *>(extra + ps.off_ybar);
    volatile int *zdone = reinterpret_cast<volatile int *>(extra + ps.off_ctrl);
    uint32_t *zbuf = reinterpret_cast<uint32_t *>(extra + ps.off_zero);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nfma = 3 * kFmaPerSmsp + a.extra_fma;
    const int cw = fma_warp_index(warp, a.extra_fma);
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
            mbar_init(&ybar[kSlabs + s], nfma);  // yfree: every FMA warp arrives once per tile
        }
        *zdone = 0;
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

    if (warp == kDpWarp) {
        // ======================= DP warp =======================
        int g = 0, k = 0;
        for (int u = blockIdx.x; u < a.B; u += gridDim.x, ++k) {
            int tx, ty, ntiles;
            bool degenerate;
            geometry(u, tx, ty, ntiles, degenerate);
            uint32_t *bits = L.bits_in_smem ? bits_smem : a.bits_ws + (size_t)u * L.nch * L.xrows;
            for (int x = lane; x < T_x; x += 32) dur[x] = 0;
            __syncwarp();
            float score = 0.0f;
            if (ntiles > 0) {
                score = prior_forward_dispatch<XPLMAX>(ring, bits, L.xrows, tx, ty, lane, g);
                g += ntiles;
                __syncwarp();
                if (lane == 0) backtrack_bits(bits, L.xrows, tx, ty, first, dur, L.bits_in_smem != 0);
            } else if (degenerate) {
                if (lane == 0) {  // reference semantics for t_x > t_y: raw prior values (mas_dp.cuh)
                    const float *mub = a.mu_x + (int64_t)u * F * T_x;
                    const float *yb = a.y + (int64_t)u * F * T_y;
                    auto val = [&](int x, int y) { return lp_cell(mub, yb, F, T_x, T_y, x, y, cst); };
                    backtrack_degenerate(val, tx, ty, first, dur);
                    score = val(tx - 1, ty - 1);
                }
                score = __shfl_sync(kFull, score, 0);
            }
            __syncwarp();
            if (lane == 0 && a.score) a.score[u] = score;
            // the output path of utterance k must have been cleared before its 1-cells are written
            if (a.path) {
                while (*zdone <= k) __nanosleep(64);
                __threadfence_block();
            }
            char *pb = a.path ? static_cast<char *>(a.path) + (int64_t)u * T_x * T_y * a.path_esize : nullptr;
            write_path_ones(pb, a.durations ? a.durations + (int64_t)u * T_x : nullptr, first, dur, T_x,
                            T_y, a.path_esize, a.one, lane, 32);
            write_frame_idx(a.frame_idx ? a.frame_idx + (int64_t)u * T_y : nullptr, first, dur, T_x, ty,
                            a.T_y, lane, 32);
            __syncwarp();
        }
    } else if (warp == kLoaderWarp) {
        // ======================= slab loader =======================
        // y[:, 32t..32t+31] -> shared (cp.async, completion signalled straight to the slab barrier),
        // and the zero-fill of the dense output path (bulk stores drained by the copy engine).
        const bool vec16 = (T_y % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.y) & 15) == 0);
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
                cp_async_arrive(&ybar[s]);  // fires when this lane's copies have landed
            }
            // zeros of utterance k are in global memory -> the DP warp may write its 1-cells
            if (zbulk) bulk_wait_all();
            __threadfence_block();
            __syncwarp();
            if (lane == 0) *zdone = k + 1;
        }
    } else if (cw >= 0) {
        // ======================= FMA warps =======================
        // Work items (tile, 32-token pass, 16-frame half) are dealt round-robin over the FMA warps,
        // continuing across tiles and utterances.  Every FMA warp waits for every slab and arrives
        // on every tile's `full` barrier (with or without work in it): parity waits are only sound
        // if no waiter can fall two phases behind or run a phase ahead of a barrier.
        const int nft = nfma * 32;
        const int ftid = cw * 32 + lane;
        int g = 0;
        int item0 = 0;  // running item counter: keeps the deal balanced across tiles/utterances
        for (int u = blockIdx.x; u < a.B; u += gridDim.x) {
            int tx, ty, ntiles;
            bool degenerate;
            geometry(u, tx, ty, ntiles, degenerate);
            if (ntiles == 0) continue;
            const int npass = (tx + 31) >> 5;
            const int xr = npass * 32;  // token rows any pass may touch
            const float *mub = a.mu_x + (int64_t)u * F * T_x;
            const RowMap rm(tx);
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
            const int nit = 2 * npass;  // items per tile
            for (int t = 0; t < ntiles; ++t, ++g) {
                const int s = g % NS, ys = g % kSlabs;
                const int lo = max(0, tx + t * kTileY - ty);
                const int hi = min(tx - 1, t * kTileY + kTileY - 1);
                if (g >= NS) mbar_wait(&ring.empty[s], ((g / NS) - 1) & 1);  // DP consumed tile g-NS
                mbar_wait(&ybar[ys], (g / kSlabs) & 1);                       // slab has landed
                int it = (cw - item0) % nfma;
                if (it < 0) it += nfma;
                for (; it < nit; it += nfma) {
                    const int p = it >> 1, h = it & 1;
                    if (32 * p <= hi && 32 * p + 31 >= lo && t * kTileY + 16 * h < ty)
                        prior_item(mu_s, musq, yslab + (size_t)ys * F * kTileY,
                                   stages + (size_t)s * ring.stage_floats, F, L.xrows, p, h, lane, cst, rm);
                }
                item0 = (item0 + nit) % nfma;
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&ring.full[s]);           // tile g produced (this warp's share)
                    mbar_arrive(&ybar[kSlabs + ys]);      // slab g no longer needed by this warp
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// unfused prior: lp[b,x,y] for every cell (same arithmetic order as the fused producers)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) log_prior_kernel(const float *__restrict__ mu_x,
                                                        const float *__restrict__ y, float *lp,
                                                        int F, int T_x, int T_y)
{
    // block: 8 token rows x 32 frames; thread (r, s) accumulates one cell
    const int b = blockIdx.z;
    const int s = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int yy = blockIdx.x * 32 + s, x = blockIdx.y * 8 + r;
    if (x >= T_x) return;
    const float *mub = mu_x + (int64_t)b * F * T_x;
    const float *yb = y + (int64_t)b * F * T_y;
    float ysq = 0.0f, c = 0.0f, msq = 0.0f;
    for (int f = 0; f < F; ++f) {
        const float m = __ldg(mub + (int64_t)f * T_x + x);
        const float v = (yy < T_y) ? __ldg(yb + (int64_t)f * T_y + yy) : 0.0f;
        ysq = __fmaf_rn(v, v, ysq);
        c = __fmaf_rn(m, v, c);
        msq = __fmaf_rn(m, m, msq);
    }
    const float cst = (float)(-0.5 * 1.8378770664093453 * (double)F);
    if (yy < T_y) lp[((int64_t)b * T_x + x) * T_y + yy] = ((-0.5f * ysq + c) + -0.5f * msq) + cst;
}

static int sm_count()
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
    const int grid = std::min(a.B, sm_count() * per_sm);
    k<<<grid, kPriorThreads, a.lay.total, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_log_prior(const float *mu_x, const float *y, float *lp, int B, int F, int T_x,
                             int T_y, cudaStream_t st)
{
    dim3 grid((T_y + 31) / 32, (T_x + 7) / 8, B);
    log_prior_kernel<<<grid, 256, 0, st>>>(mu_x, y, lp, F, T_x, T_y);
    count_launch();
    return cudaGetLastError();
}

}  // namespace mas
