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

constexpr int kPriorThreads = 288;                     // 1 DP warp + 8 producer warps
constexpr int kProducerWarps = kPriorThreads / 32 - 1;

__host__ __device__ inline int mu_pitch(int F)
{
    int p = (F + 3) & ~3;          // float4 reads along f
    if (((p >> 2) & 1) == 0) p += 4;  // odd number of 16-byte chunks per row: spreads banks
    return p;
}

struct PriorSmem {
    size_t off_mu, off_musq, off_ytile, total_extra;
};

__host__ __device__ inline PriorSmem prior_smem(int F, int xrows)
{
    PriorSmem s;
    const int P = mu_pitch(F);
    s.off_mu = 0;
    s.off_musq = s.off_mu + (size_t)xrows * P * 4;
    s.off_ytile = s.off_musq + (size_t)xrows * 4;
    s.total_extra = s.off_ytile + 2 * (size_t)P * kTileY * 4;
    s.total_extra = (s.total_extra + 15) & ~(size_t)15;
    return s;
}

size_t prior_extra_smem(int F, int T_x)
{
    return prior_smem(F, ((T_x + 31) / 32) * 32).total_extra;
}

__device__ __forceinline__ void producer_bar()
{
    asm volatile("bar.sync 1, %0;" ::"n"(kProducerWarps * 32) : "memory");
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

template <int XPLMAX>
__global__ void __launch_bounds__(kPriorThreads) mas_prior_kernel(const PriorArgs a)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const FastLayout &L = a.lay;
    const int F = a.F, T_x = a.T_x;
    const int64_t T_y = a.T_y;
    const int P = mu_pitch(F);
    const PriorSmem ps = prior_smem(F, L.xrows);
    float *stages = reinterpret_cast<float *>(smem + L.off_stages);
    int *first = reinterpret_cast<int *>(smem + L.off_first);
    int *dur = reinterpret_cast<int *>(smem + L.off_dur);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.off_bars);
    unsigned char *extra = smem + L.off_bars + 128;
    float *mu_s = reinterpret_cast<float *>(extra + ps.off_mu);      // [xrows][P]
    float *musq = reinterpret_cast<float *>(extra + ps.off_musq);    // [xrows]
    float *ytile = reinterpret_cast<float *>(extra + ps.off_ytile);  // [2][32 frames][P]

    const int b = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tx = min(max(a.t_x[b], 0), T_x);
    const int ty = min(max(a.t_y[b], 0), a.T_y);
    const bool degenerate = tx > ty && ty >= 1;
    const bool active = tx >= 1 && ty >= 1 && !degenerate;
    const int ntiles = active ? (ty + kTileY - 1) / kTileY : 0;

    uint32_t *bits = L.bits_in_smem ? reinterpret_cast<uint32_t *>(smem + L.off_bits)
                                    : a.bits_ws + (size_t)b * L.nch * L.xrows;
    TileRing ring;
    ring.stages = stages;
    ring.full = bars;
    ring.empty = bars + L.nstages;
    ring.nstages = L.nstages;
    ring.stage_floats = L.xrows * kTileY;
    if (tid == 0) {
        for (int s = 0; s < L.nstages; ++s) {
            mbar_init(&ring.full[s], kProducerWarps);
            mbar_init(&ring.empty[s], 1);
        }
        mbar_fence_init();
    }

    // mu_x[b] -> shared, transposed to [token][feature] so that a producer reads four
    // features of one token with a single broadcast LDS.128.
    const float *mub = a.mu_x + (int64_t)b * F * T_x;
    const float *yb = a.y + (int64_t)b * F * T_y;
    if (active) {
        for (int i = tid; i < F * tx; i += kPriorThreads) {
            const int f = i / tx, x = i - f * tx;
            mu_s[x * P + f] = __ldg(mub + (int64_t)f * T_x + x);
        }
        for (int i = tid; i < (P - F) * tx; i += kPriorThreads) {  // zero the pad features
            const int x = i / (P - F), f = F + i - x * (P - F);
            mu_s[x * P + f] = 0.0f;
        }
    }
    __syncthreads();
    if (active) {
        for (int x = tid; x < tx; x += kPriorThreads) {
            float s = 0.0f;
            for (int f = 0; f < F; ++f) {
                const float m = mu_s[x * P + f];
                s = __fmaf_rn(m, m, s);
            }
            musq[x] = -0.5f * s;  // tts.py:494  mu_square = sum(factor * mu^2)
        }
    }
    __syncthreads();

    const float cst = (float)(-0.5 * 1.8378770664093453 * (double)F);  // -0.5*log(2*pi)*F, tts.py:484
    char *pb = a.path ? static_cast<char *>(a.path) + (int64_t)b * T_x * T_y * a.path_esize : nullptr;
    const int64_t pbytes = a.path ? (int64_t)T_x * T_y * a.path_esize : 0;

    if (warp == 0) {
        for (int x = lane; x < T_x; x += 32) dur[x] = 0;
        __syncwarp();
        float score = 0.0f;
        if (active) {
            score = prior_forward_dispatch<XPLMAX>(ring, bits, L.xrows, tx, ty, lane);
            __syncwarp();
            if (lane == 0) backtrack_bits(bits, L.xrows, tx, ty, first, dur);
        } else if (degenerate) {
            if (lane == 0) {  // reference semantics for t_x > t_y: raw prior values, see mas_dp.cuh
                auto val = [&](int x, int y) { return lp_cell(mub, yb, F, T_x, T_y, x, y, cst); };
                backtrack_degenerate(val, tx, ty, first, dur);
                score = val(tx - 1, ty - 1);
            }
            score = __shfl_sync(kFull, score, 0);
        }
        if (lane == 0 && a.score) a.score[b] = score;
    } else {
        const int pw = warp - 1;
        const int ptid = tid - 32;
        constexpr int npt = kProducerWarps * 32;
        int stage = 0;
        uint32_t phase = 0;
        // prologue: y slab of tile 0
        auto load_y = [&](int t, float *dst) {
            // dst[frame][f]; global reads coalesced along frames, one feature row at a time
            const int y0 = t * kTileY;
            for (int i = ptid; i < F * kTileY; i += npt) {
                const int f = i >> 5, s = i & 31;
                const int y = y0 + s;
                dst[s * P + f] = (y < ty) ? __ldg(yb + (int64_t)f * T_y + y) : 0.0f;
            }
            for (int i = ptid; i < (P - F) * kTileY; i += npt) {
                const int s = i / (P - F), f = F + i - s * (P - F);
                dst[s * P + f] = 0.0f;
            }
        };
        if (ntiles > 0) load_y(0, ytile);
        for (int t = 0; t < ntiles; ++t) {
            producer_bar();  // slab t visible to all producers; slab t-1 no longer read
            if (t + 1 < ntiles) load_y(t + 1, ytile + ((t + 1) & 1) * P * kTileY);
            if (t >= L.nstages) mbar_wait(&ring.empty[stage], phase ^ 1u);
            float *tile = stages + stage * ring.stage_floats;
            const float *ys = ytile + (t & 1) * P * kTileY + lane * P;  // this lane's frame
            // -0.5 * |y_j|^2 for this lane's frame (tts.py:488-490)
            float ysq = 0.0f;
            for (int f = 0; f < P; f += 4) {
                const float4 v = *reinterpret_cast<const float4 *>(ys + f);
                ysq = __fmaf_rn(v.x, v.x, ysq);
                ysq = __fmaf_rn(v.y, v.y, ysq);
                ysq = __fmaf_rn(v.z, v.z, ysq);
                ysq = __fmaf_rn(v.w, v.w, ysq);
            }
            ysq *= -0.5f;
            const int lo = max(0, tx + t * kTileY - ty);
            const int hi = min(tx - 1, t * kTileY + kTileY - 1);
            // two token rows per pass: every y value read from shared memory feeds two FMAs
            for (int x = lo + 2 * pw; x <= hi; x += 2 * kProducerWarps) {
                const int x1 = min(x + 1, hi);
                const float *m0 = mu_s + x * P, *m1 = mu_s + x1 * P;
                float c0 = 0.0f, c1 = 0.0f;
#pragma unroll 4
                for (int f = 0; f < P; f += 4) {
                    const float4 yv = *reinterpret_cast<const float4 *>(ys + f);
                    const float4 a0 = *reinterpret_cast<const float4 *>(m0 + f);
                    const float4 a1 = *reinterpret_cast<const float4 *>(m1 + f);
                    c0 = __fmaf_rn(a0.x, yv.x, c0);
                    c1 = __fmaf_rn(a1.x, yv.x, c1);
                    c0 = __fmaf_rn(a0.y, yv.y, c0);
                    c1 = __fmaf_rn(a1.y, yv.y, c1);
                    c0 = __fmaf_rn(a0.z, yv.z, c0);
                    c1 = __fmaf_rn(a1.z, yv.z, c1);
                    c0 = __fmaf_rn(a0.w, yv.w, c0);
                    c1 = __fmaf_rn(a1.w, yv.w, c1);
                }
                // tts.py:495: y_square - y_mu_double + mu_square + const, y_mu_double = -cross
                const float lp0 = ((ysq + c0) + musq[x]) + cst;
                const float lp1 = ((ysq + c1) + musq[x1]) + cst;
                tile[tile_index(x, lane)] = lp0;
                if (x1 != x) tile[tile_index(x1, lane)] = lp1;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ring.full[stage]);
            if (++stage == L.nstages) {
                stage = 0;
                phase ^= 1u;
            }
            zero_fill_part(pb, pbytes, t, ntiles, ptid, npt);
        }
        if (ntiles == 0) zero_fill_part(pb, pbytes, 0, 1, ptid, npt);
    }
    __syncthreads();
    write_path_ones(pb, a.durations ? a.durations + (int64_t)b * T_x : nullptr, first, dur, T_x, T_y,
                    a.path_esize, a.one, tid, kPriorThreads);
    write_frame_idx(a.frame_idx ? a.frame_idx + (int64_t)b * T_y : nullptr, first, dur, T_x, ty,
                    a.T_y, tid, kPriorThreads);
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
