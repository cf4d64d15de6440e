// mas_prior.cu -- fused Gaussian log-prior + MAS for sm_100a.
//
// Replaces the block GradTTS/ArtTTS.compute_loss runs under torch.no_grad()
// (src/model/tts.py:483-500 and its copies at :200-214, :776-790, :1067-1081) plus the
// duration sum of tts.py:503-505:
//
//     lp[x,j] = -0.5*sum_f y[f,j]^2 + sum_f mu[f,x]*y[f,j] - 0.5*sum_f mu[f,x]^2 - 0.5*F*log(2*pi)
//     path    = maximum_path(lp, mask);   durations = sum_j path
//
// mas_prior_kernel  one CTA per utterance.  mu_x of the utterance lives in shared memory for
//                   the whole CTA lifetime; producer warps turn 32-frame slabs of y into
//                   32-frame log-prior tiles (fp32 FMA on CUDA cores: with F <= 80 the
//                   contraction is far too thin for tensor cores) written straight into the
//                   same swizzled ring the drop-in kernel fills from HBM, and warp 0 consumes
//                   them with the identical single-warp recurrence (mas_dp.cuh).  The
//                   T_x x T_y matrix never touches HBM; only the band of each tile is computed.
// log_prior_kernel  the unfused prior (parity tap `log_prior_out`, and the fallback for shapes
//                   whose operands do not fit in shared memory).
#include <cmath>

#include "mas_dp.cuh"
#include "mas_internal.h"

namespace mas {

// Optional phase timing (profiles/prior_timeline.py builds a -DMAS_TIMING copy of the library;
// the production build compiles all of it away).
#ifdef MAS_TIMING
#define MAS_T_DECL long long t_acc0 = 0, t_acc1 = 0, t_acc2 = 0, t_acc3 = 0, t_tmp = 0; (void)t_acc0; (void)t_acc1; (void)t_acc2; (void)t_acc3; (void)t_tmp
#define MAS_T_BEGIN() (t_tmp = clock64())
#define MAS_T_END(acc) ((acc) += clock64() - t_tmp)
#define MAS_T_PUT(i, v) do { if (a.timing && lane == 0) a.timing[(size_t)b * 32 + (i)] = (v); } while (0)
#else
#define MAS_T_DECL
#define MAS_T_BEGIN()
#define MAS_T_END(acc)
#define MAS_T_PUT(i, v)
#endif

// Warp roles.  A warp's scheduler (SMSP) is warp_id % 4.  The frame-sequential DP warp needs
// most of one scheduler's issue slots to run at its natural ~85 cycles/frame, and it cannot be
// given priority over the FMA warps' long independent streams (measured: sharing a scheduler
// with 4 FMA warps serialised the two).  So scheduler 3 is reserved for the latency-bound
// agents -- DP warp (warp 3) and slab loader (warp 7) -- plus an optional, tunable number of
// FMA warps (warps 11, 15); schedulers 0-2 run 4 FMA warps each.
constexpr int kFmaPerSmsp = 4;
constexpr int kPriorWarps = 16;
constexpr int kPriorThreads = 32 * kPriorWarps;
constexpr int kMaxFmaWarps = 3 * kFmaPerSmsp + 2;
constexpr int kDpWarp = 3;
constexpr int kLoaderWarp = 7;
constexpr int kZeroBytes = 8192;  // zeroed shared buffer behind the bulk stores

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
//   ybar  [kSlabs]          slab-ready mbarriers
//   zero  [kZeroBytes]      zeros, source of the bulk stores that clear the output path
struct PriorSmem {
    size_t off_mu, off_musq, off_yslab, off_ybar, off_zero, total_extra;
};

constexpr int kSlabs = 4;

__host__ __device__ inline PriorSmem prior_smem(int F, int xrows)
{
    PriorSmem s;
    s.off_mu = 0;
    s.off_musq = s.off_mu + (size_t)F * xrows * 4;
    s.off_yslab = s.off_musq + (size_t)xrows * 4;
    s.off_ybar = s.off_yslab + (size_t)kSlabs * F * kTileY * 4;
    s.off_zero = (s.off_ybar + (size_t)kSlabs * 8 + 15) & ~(size_t)15;
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
                                                        int xrows, int tx, int ty, int lane)
{
    const int xpl = (tx + 31) >> 5;
#define MAS_CASE(N)                                                                       \
    case N:                                                                               \
        if constexpr (N <= XPLMAX) return dp_forward<N>(ring, bits, xrows, tx, ty, lane); \
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

// 32 tokens x 32 frames of the prior by one warp: thread (xg, yg) owns 4 tokens x 8 frames,
// i.e. 32 independent fp32 FMA chains fed by 3 LDS.128 per feature (one conflict-free read of
// mu, two of y).  Accumulation runs over f in ascending order with one FMA per term, exactly
// like log_prior_kernel / lp_cell, so all three produce bit-identical values.
__device__ __forceinline__ void prior_pass(const float *__restrict__ mu_s, const float *__restrict__ musq,
                                           const float *__restrict__ ys, float *__restrict__ tile, int F,
                                           int xrows, int p, int lane, float cst, const RowMap rm)
{
    const int xg = lane >> 2, yg = lane & 3;
    const int x0 = 32 * p + 4 * xg;
    const float *mp = mu_s + x0;
    const float *yp = ys + 8 * yg;
    // -0.5*|y_j|^2 (tts.py:488-490): the 8 lanes that share this thread's 8 frames (same yg)
    // each accumulate ONE of them (frame 8*yg + xg) alongside the main loop -- one extra LDS +
    // FMA per feature -- and the values are exchanged with shuffles at the end.
    const float *yq = ys + 8 * yg + xg;
    float qsum = 0.0f;
    float acc[4][8];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[r][k] = 0.0f;
#pragma unroll 2
    for (int f = 0; f < F; ++f) {
        const float4 m = *reinterpret_cast<const float4 *>(mp + (size_t)f * xrows);
        const float4 ya = *reinterpret_cast<const float4 *>(yp + f * kTileY);
        const float4 yb = *reinterpret_cast<const float4 *>(yp + f * kTileY + 4);
        const float yo = yq[f * kTileY];
        const float mr[4] = {m.x, m.y, m.z, m.w};
        const float yk[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[r][k] = __fmaf_rn(mr[r], yk[k], acc[r][k]);
        qsum = __fmaf_rn(yo, yo, qsum);
    }
    qsum *= -0.5f;
    float q[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) q[k] = __shfl_sync(kFull, qsum, (k << 2) | yg);  // lane (xg=k, yg)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int x = x0 + r;
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
    uint32_t *zbuf = reinterpret_cast<uint32_t *>(extra + ps.off_zero);

    const int b = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tx = min(max(a.t_x[b], 0), T_x);
    const int ty = min(max(a.t_y[b], 0), a.T_y);
    const bool degenerate = tx > ty && ty >= 1;
    const bool active = tx >= 1 && ty >= 1 && !degenerate;
    const int ntiles = active ? (ty + kTileY - 1) / kTileY : 0;
    const int npass = (tx + 31) >> 5;  // 32-token passes per tile
    const int nfma = 3 * kFmaPerSmsp + a.extra_fma;

    uint32_t *bits = L.bits_in_smem ? reinterpret_cast<uint32_t *>(smem + L.off_bits)
                                    : a.bits_ws + (size_t)b * L.nch * L.xrows;
    MAS_T_DECL;
    if (warp == 0) MAS_T_PUT(0, clock64());
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
        for (int s = 0; s < kSlabs; ++s) mbar_init(&ybar[s], 32);  // one cp.async arrival per loader lane
        mbar_fence_init();
    }

    const float *mub = a.mu_x + (int64_t)b * F * T_x;
    const float *yb = a.y + (int64_t)b * F * T_y;
    const int xr = npass * 32;  // token rows any pass may touch
    for (int i = tid; i < kZeroBytes / 4; i += kPriorThreads) zbuf[i] = 0u;
    fence_proxy_async_smem();  // generic-proxy zeros -> visible to the bulk-copy (async) proxy
    if (active) {
        for (int i = tid; i < F * xr; i += kPriorThreads) {  // all requests in flight at once
            const int f = i / xr, x = i - f * xr;
            cp_async4(mu_s + f * L.xrows + x, mub + (int64_t)f * T_x + (x < tx ? x : 0),
                      x < tx ? 4u : 0u);
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    __syncthreads();
    if (active) {
        for (int x = tid; x < xr; x += kPriorThreads) {
            float s = 0.0f;
            int f = 0;
            for (; f + 16 <= F; f += 16) {
                float m[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) m[u] = mu_s[(f + u) * L.xrows + x];
#pragma unroll
                for (int u = 0; u < 16; ++u) s = __fmaf_rn(m[u], m[u], s);
            }
            for (; f < F; ++f) {
                const float m = mu_s[f * L.xrows + x];
                s = __fmaf_rn(m, m, s);
            }
            musq[x] = -0.5f * s;  // tts.py:494  mu_square = sum(factor * mu^2)
        }
    }
    __syncthreads();

    if (warp == 0) MAS_T_PUT(1, clock64());
    const float cst = (float)(-0.5 * 1.8378770664093453 * (double)F);  // -0.5*log(2*pi)*F, tts.py:484
    char *pb = a.path ? static_cast<char *>(a.path) + (int64_t)b * T_x * T_y * a.path_esize : nullptr;
    const int64_t pbytes = a.path ? (int64_t)T_x * T_y * a.path_esize : 0;

    if (warp == kDpWarp) {
        // ---------------- DP warp ----------------
        for (int x = lane; x < T_x; x += 32) dur[x] = 0;
        __syncwarp();
        float score = 0.0f;
        if (active) {
            score = prior_forward_dispatch<XPLMAX>(ring, bits, L.xrows, tx, ty, lane);
            __syncwarp();
            MAS_T_PUT(10, clock64());
            if (lane == 0) backtrack_bits(bits, L.xrows, tx, ty, first, dur, L.bits_in_smem != 0);
            MAS_T_PUT(11, clock64());
        } else if (degenerate) {
            if (lane == 0) {  // reference semantics for t_x > t_y: raw prior values, see mas_dp.cuh
                auto val = [&](int x, int y) { return lp_cell(mub, yb, F, T_x, T_y, x, y, cst); };
                backtrack_degenerate(val, tx, ty, first, dur);
                score = val(tx - 1, ty - 1);
            }
            score = __shfl_sync(kFull, score, 0);
        }
        if (lane == 0 && a.score) a.score[b] = score;
    } else if (warp == kLoaderWarp) {
        // ---------------- slab loader: y[:, 32t..32t+31] -> shared (cp.async, completion signalled
        // straight to the slab barrier), and the zero-fill of the dense output path
        const bool vec16 = (T_y % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.y) & 15) == 0);
        const bool zbulk = bulk_zero_ok(pb, pbytes);
        auto zero_part = [&](int part, int nparts) {
            if (zbulk) zero_fill_bulk_part(pb, pbytes, part, nparts, zbuf, kZeroBytes, lane, 32);
            else zero_fill_part(pb, pbytes, part, nparts, lane, 32);
        };
        auto issue = [&](int t) {
            const int s = t % kSlabs;
            if (t >= kSlabs) {  // slot's previous slab (tile t-kSlabs) fully used by the FMA warps
                const int u = t - kSlabs;
                MAS_T_BEGIN();
                mbar_wait(&ring.full[u % NS], (u / NS) & 1);
                MAS_T_END(t_acc1);
            }
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
                for (int f = 0; f < F; ++f) cp_async4(dst + f * kTileY + lane, src + (int64_t)f * T_y, bytes);
            }
            cp_async_arrive(&ybar[s]);  // slab-ready barrier fires when this lane's copies have landed
        };
        // the dense output is cleared up front: ~80 bulk stores handed to the copy engine, which
        // drains them in the background while the tiles are produced
        zero_part(0, 1);
        // y runs up to kSlabs-1 tiles ahead of the arithmetic; the loader depends only on the FMA
        // warps (slot free <=> tile t-kSlabs fully produced), never on the DP warp.
        constexpr int depth = kSlabs - 1;
        for (int t = 0; t < min(depth, ntiles); ++t) issue(t);
        for (int t = 0; t + depth < ntiles; ++t) issue(t + depth);
        if (zbulk) {  // the zeros must be in global memory before any thread writes a 1 over them
            bulk_commit();
            bulk_wait_all();
        }
        MAS_T_PUT(5, clock64());
        MAS_T_PUT(7, t_acc1);
    } else if (fma_warp_index(warp, a.extra_fma) >= 0) {
        // ---------------- FMA warps: work items (tile t, pass p), dealt round-robin ----------------
        // Every FMA warp waits for every slab and arrives on every tile's `full` barrier (with or
        // without work in it): parity waits are only sound if no waiter can fall two phases behind
        // or run a phase ahead of a barrier, and this makes both impossible by construction.
        const int cw = fma_warp_index(warp, a.extra_fma);
        const RowMap rm(tx);
        int s = 0;
        uint32_t phase = 0;
        for (int t = 0; t < ntiles; ++t) {
            const int lo = max(0, tx + t * kTileY - ty);
            const int hi = min(tx - 1, t * kTileY + kTileY - 1);
            const int ys = t % kSlabs;
            MAS_T_BEGIN();
            if (t >= NS) mbar_wait(&ring.empty[s], phase ^ 1u);  // DP warp consumed tile t-NS
            MAS_T_END(t_acc0);
            MAS_T_BEGIN();
            mbar_wait(&ybar[ys], (t / kSlabs) & 1);               // slab t has landed in shared memory
            MAS_T_END(t_acc1);
            int p = (cw - t * npass) % nfma;
            if (p < 0) p += nfma;
            for (; p < npass; p += nfma)
                if (32 * p <= hi && 32 * p + 31 >= lo)
                    prior_pass(mu_s, musq, yslab + (size_t)ys * F * kTileY,
                               stages + (size_t)s * ring.stage_floats, F, L.xrows, p, lane, cst, rm);
            __syncwarp();
            if (lane == 0) mbar_arrive(&ring.full[s]);
            if (++s == NS) {
                s = 0;
                phase ^= 1u;
            }
        }
        if (cw == 0) {
            MAS_T_PUT(2, clock64());
            MAS_T_PUT(3, t_acc0);
            MAS_T_PUT(4, t_acc1);
        }
    }
    __syncthreads();
    write_path_ones(pb, a.durations ? a.durations + (int64_t)b * T_x : nullptr, first, dur, T_x, T_y,
                    a.path_esize, a.one, tid, kPriorThreads);
    write_frame_idx(a.frame_idx ? a.frame_idx + (int64_t)b * T_y : nullptr, first, dur, T_x, ty,
                    a.T_y, tid, kPriorThreads);
    if (warp == 0) MAS_T_PUT(12, clock64());
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
    k<<<a.B, kPriorThreads, a.lay.total, st>>>(a);
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
