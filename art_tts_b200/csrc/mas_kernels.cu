// mas_kernels.cu -- drop-in maximum_path kernels for sm_100a.
//
// Replaces maximum_path_c (src/model/monotonic_align/core.pyx:38-45) and the host glue of
// maximum_path (src/model/monotonic_align/__init__.py:13-23) of antoinelii/art-tts.
//
// mas_fast_kernel   one CTA per utterance: warp 0 runs the frame-sequential recurrence on
//                   32-frame tiles that four staging warps stream from HBM into a swizzled
//                   shared-memory ring (coalesced along frames, band-limited), while the
//                   same staging warps zero-fill the dense output path.  Direction bits
//                   stay in shared memory (or spill to the workspace for long utterances);
//                   the backtrack runs on them and the whole CTA then writes the 1-cells.
//                   The fp32 score matrix is read exactly once and never written.
// mas_general_kernel  size-agnostic fallback (T_x > 512): block-wide row sweep, one
//                   __syncthreads per frame, direction bits in the workspace.
#include "mas_dp.cuh"
#include "mas_internal.h"

namespace mas {

// ------------------------------------------------------------------------------------
// lengths from the mask (monotonic_align/__init__.py:18-21)
// ------------------------------------------------------------------------------------
__device__ __forceinline__ double mask_elem(const void *m, int dtype, int64_t i)
{
    switch (dtype) {
    case MAS_F32: return static_cast<const float *>(m)[i];
    case MAS_F16: return __half2float(static_cast<const __half *>(m)[i]);
    case MAS_BF16: return __bfloat162float(static_cast<const __nv_bfloat16 *>(m)[i]);
    case MAS_F64: return static_cast<const double *>(m)[i];
    case MAS_I32: return static_cast<const int32_t *>(m)[i];
    case MAS_U8: return static_cast<const uint8_t *>(m)[i];
    default: return static_cast<double>(static_cast<const int64_t *>(m)[i]);
    }
}

__global__ void __launch_bounds__(128) lengths_kernel(const void *mask, int dtype, int T_x, int T_y,
                                                      int64_t sb, int64_t sx, int64_t sy,
                                                      int32_t *t_x, int32_t *t_y)
{
    const int b = blockIdx.x;
    double ax = 0.0, ay = 0.0;
    for (int x = threadIdx.x; x < T_x; x += blockDim.x) ax += mask_elem(mask, dtype, b * sb + x * sx);
    for (int y = threadIdx.x; y < T_y; y += blockDim.x) ay += mask_elem(mask, dtype, b * sb + y * sy);
    __shared__ double red[2][4];
    for (int o = 16; o > 0; o >>= 1) {
        ax += __shfl_xor_sync(kFull, ax, o);
        ay += __shfl_xor_sync(kFull, ay, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = ax;
        red[1][threadIdx.x >> 5] = ay;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        t_x[b] = static_cast<int32_t>(red[0][0] + red[0][1] + red[0][2] + red[0][3]);
        t_y[b] = static_cast<int32_t>(red[1][0] + red[1][1] + red[1][2] + red[1][3]);
    }
}

cudaError_t launch_lengths_from_mask(const void *mask, int mask_dtype, int B, int T_x, int T_y,
                                     int64_t sb, int64_t sx, int64_t sy, int32_t *t_x, int32_t *t_y,
                                     cudaStream_t st)
{
    lengths_kernel<<<B, 128, 0, st>>>(mask, mask_dtype, T_x, T_y, sb, sx, sy, t_x, t_y);
    count_launch();
    return cudaGetLastError();
}

// lengths from the two sequence masks of the fused entry (tts.py:477-480 builds attn_mask from
// x_mask [B,1,T_x] and y_mask [B,1,T_y]; summing them is what mask.sum(1)[:,0] / mask.sum(2)[:,0]
// of their outer product gives, __init__.py:20-21): one launch instead of four eager torch ops
__global__ void __launch_bounds__(128) seq_lengths_kernel(const void *xm, int xdt, int64_t xsb, int64_t xst,
                                                          const void *ym, int ydt, int64_t ysb, int64_t yst,
                                                          int T_x, int T_y, int32_t *t_x, int32_t *t_y)
{
    const int b = blockIdx.x;
    double ax = 0.0, ay = 0.0;
    for (int x = threadIdx.x; x < T_x; x += blockDim.x) ax += mask_elem(xm, xdt, b * xsb + x * xst);
    for (int y = threadIdx.x; y < T_y; y += blockDim.x) ay += mask_elem(ym, ydt, b * ysb + y * yst);
    __shared__ double red[2][4];
    for (int o = 16; o > 0; o >>= 1) {
        ax += __shfl_xor_sync(kFull, ax, o);
        ay += __shfl_xor_sync(kFull, ay, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = ax;
        red[1][threadIdx.x >> 5] = ay;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        t_x[b] = static_cast<int32_t>(red[0][0] + red[0][1] + red[0][2] + red[0][3]);
        t_y[b] = static_cast<int32_t>(red[1][0] + red[1][1] + red[1][2] + red[1][3]);
    }
}

cudaError_t launch_seq_lengths(const void *xm, int xdt, int64_t xsb, int64_t xst, const void *ym, int ydt,
                               int64_t ysb, int64_t yst, int B, int T_x, int T_y, int32_t *t_x, int32_t *t_y,
                               cudaStream_t st)
{
    seq_lengths_kernel<<<B, 128, 0, st>>>(xm, xdt, xsb, xst, ym, ydt, ysb, yst, T_x, T_y, t_x, t_y);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// fast kernel
// ------------------------------------------------------------------------------------
// Stage the band of tile t (frames 32t..32t+31, tokens lo..hi) into shared memory.
// One warp per token row, lanes along frames: every global request is one contiguous
// 128-byte segment, every shared store is conflict-free under the XOR swizzle.
template <typename InT>
__device__ __forceinline__ void stage_value_tile(const InT *__restrict__ vb,
                                                 const float *__restrict__ mb, float *stage, int t,
                                                 int tx, int ty, int64_t T_y, int hw, int nhw,
                                                 int lane)
{
    const RowMap rm(tx);
    const int y = t * kTileY + lane;
    const bool in = y < ty;
    const int lo = max(0, tx + t * kTileY - ty);
    const int hi = min(tx - 1, t * kTileY + kTileY - 1);
    constexpr int U = 8;
    for (int x0 = lo + hw; x0 <= hi; x0 += U * nhw) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int x = x0 + u * nhw;
            v[u] = 0.0f;
            if (x <= hi && in) {
                const int64_t e = (int64_t)x * T_y + y;
                v[u] = load_as_f32(vb + e);
                if (mb) v[u] *= __ldg(mb + e);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int x = x0 + u * nhw;
            if (x <= hi) stage[tile_index(rm.row(x), lane)] = v[u];
        }
    }
}

// fp32, unmasked: the same band through cp.async -- nothing waits on a register, so a staging
// thread can have every row of several tiles in flight (HBM latency is hidden by depth, not
// by occupancy).  VEC = frames per request: 4 when rows are 16-byte aligned (T_y % 4 == 0), 2 when they are
// 8-byte aligned (T_y even: the reference's un-padded 870-frame batches), else 1.
template <int VEC, int LSH = 5>
__device__ __forceinline__ void stage_value_tile_async(const float *__restrict__ vb, float *stage,
                                                       int t, int tx, int ty, int64_t T_y, int hw,
                                                       int nhw, int lane)
{
    const RowMap rm(tx, LSH);
    const int y0 = t * kTileY;
    const int lo = max(0, tx + y0 - ty);
    const int hi = min(tx - 1, y0 + kTileY - 1);
    if constexpr (VEC == 2) {
        const int c = lane & 15, r = lane >> 4;  // 8-byte chunk of the row, row within a group of 2
        const int left = ty - (y0 + 2 * c);
        const uint32_t bytes = left >= 2 ? 8u : (left > 0 ? 4u : 0u);
        const int yo = bytes ? y0 + 2 * c : 0;
        for (int x = lo + 2 * hw + r; x <= hi; x += 2 * nhw) {
            const int row = rm.row(x);
            cp_async8(stage + (row << 5) + (((c >> 1) ^ (row & 7)) << 2) + ((c & 1) << 1), vb + (int64_t)x * T_y + yo, bytes);
        }
    } else if constexpr (VEC == 4) {
        const int c = lane & 7, r = lane >> 3;   // 16-byte chunk of the row, row within a group of 4
        const int left = ty - (y0 + 4 * c);      // valid frames from this chunk on
        const uint32_t bytes = left >= 4 ? 16u : (left > 0 ? 4u * left : 0u);
        const int yo = bytes ? y0 + 4 * c : 0;   // nothing is read when bytes == 0
        for (int x = lo + 4 * hw + r; x <= hi; x += 4 * nhw) {
            const int row = rm.row(x);
            cp_async16(stage + (row << 5) + ((c ^ (row & 7)) << 2), vb + (int64_t)x * T_y + yo, bytes);
        }
    } else {
        const int y = y0 + lane;
        const uint32_t bytes = y < ty ? 4u : 0u;
        const float *src = vb + (y < ty ? y : 0);
        for (int x = lo + hw; x <= hi; x += nhw)
            cp_async4(stage + tile_index(rm.row(x), lane), src + (int64_t)x * T_y, bytes);
    }
}

template <int XPL, bool NAT>
__device__ __forceinline__ float run_forward(const TileRing &ring, uint32_t *bits, int xrows, int tx,
                                             int ty, int lane)
{
    return dp_forward<XPL, false, NAT>(ring, bits, xrows, tx, ty, lane);
}

// dispatch on tokens-per-lane of THIS utterance (warp-uniform), bounded by the launch bucket.
// NAT: tiles in natural row order (TMA boxes), see dp_tile
template <int XPLMAX, bool NAT = false>
__device__ __forceinline__ float forward_dispatch(const TileRing &ring, uint32_t *bits, int xrows,
                                                  int tx, int ty, int lane)
{
    const int xpl = (tx + 31) >> 5;
#define MAS_CASE(N)                                                               \
    case N:                                                                       \
        if constexpr (N <= XPLMAX) return run_forward<N, NAT>(ring, bits, xrows, tx, ty, lane); \
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

template <typename InT, int XPLMAX>
__global__ void __launch_bounds__(kFastThreads) mas_fast_kernel(const MasArgs a, const __grid_constant__ TensorMap tmap)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const FastLayout &L = a.lay;
    float *stages = reinterpret_cast<float *>(smem + L.off_stages);
    int *first = reinterpret_cast<int *>(smem + L.off_first);
    int *dur = reinterpret_cast<int *>(smem + L.off_dur);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.off_bars);
    uint32_t *zbuf = reinterpret_cast<uint32_t *>(smem + L.off_bars + 128);   // kFastZeroBytes of zeros (bulk fill)

    const int b = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T_x = a.T_x;
    const int64_t T_y = a.T_y;
    const int tx = min(max(a.t_x[b], 0), T_x);
    const int ty = min(max(a.t_y[b], 0), a.T_y);
    const bool degenerate = tx > ty && ty >= 1;
    const bool active = tx >= 1 && ty >= 1 && !degenerate;
    const int ntiles = active ? (ty + kTileY - 1) / kTileY : 0;
    constexpr int kHelperWarps = kFastThreads / 32 - 1;

    uint32_t *bits = L.bits_in_smem
                         ? reinterpret_cast<uint32_t *>(smem + L.off_bits)
                         : a.bits_ws + (size_t)b * L.nch * L.xrows;
    TileRing ring;
    ring.stages = stages;
    ring.full = bars;
    ring.empty = bars + L.nstages;
    ring.nstages = L.nstages;
    ring.stage_floats = L.srows * kTileY;
    const bool tma = sizeof(InT) == 4 && a.load_mode == 3;

    if (tid == 0) {
        for (int s = 0; s < L.nstages; ++s) {
            // TMA staging: one thread announces the tile's bytes; else every staging thread arrives once
            mbar_init(&ring.full[s], tma ? 1 : kHelperWarps * 32);
            mbar_init(&ring.empty[s], 1);
        }
        mbar_fence_init();
        if (tma) tma_prefetch_desc(&tmap);
    }
    for (int i = tid; i < kFastZeroBytes / 4; i += kFastThreads) zbuf[i] = 0u;
    fence_proxy_async_smem();
    __syncthreads();

    const InT *vb = static_cast<const InT *>(a.value) + (int64_t)b * T_x * T_y;
    const float *mb = a.cell_mask ? a.cell_mask + (int64_t)b * T_x * T_y : nullptr;

    if (warp == kHelperWarps) {
        // ---------------- DP warp (highest warp id = highest issue priority on its SMSP):
        // forward recurrence + backtrack ----------------
        for (int x = lane; x < T_x; x += 32) dur[x] = 0;
        __syncwarp();
        float score = 0.0f;
        if (active) {
            if constexpr (sizeof(InT) == 4) {
                score = tma ? forward_dispatch<XPLMAX, true>(ring, bits, L.xrows, tx, ty, lane)
                            : forward_dispatch<XPLMAX, false>(ring, bits, L.xrows, tx, ty, lane);
            } else {
                score = forward_dispatch<XPLMAX, false>(ring, bits, L.xrows, tx, ty, lane);
            }
            __syncwarp();
            if (lane == 0) backtrack_bits(bits, L.xrows, tx, ty, first, dur, L.bits_in_smem != 0);
        } else if (degenerate) {
            if (lane == 0) {
                auto val = [&](int x, int y) {
                    const int64_t e = (int64_t)x * T_y + y;
                    float v = load_as_f32(vb + e);
                    if (mb) v *= mb[e];
                    return v;
                };
                backtrack_degenerate(val, tx, ty, first, dur);
                score = val(tx - 1, ty - 1);  // the reference leaves value[t_x-1,t_y-1] untouched
            }
            score = __shfl_sync(kFull, score, 0);
        }
        if (lane == 0 && a.score) a.score[b] = score;
    } else {
        // ---------------- staging warps: HBM -> ring, and zero-fill of the output ----------------
        const int hw = warp;
        const int htid = tid;
        constexpr int nht = kHelperWarps * 32;
        char *pb = a.path ? static_cast<char *>(a.path) + (int64_t)b * T_x * T_y * a.path_esize
                          : nullptr;
        const int64_t pbytes = a.path ? (int64_t)T_x * T_y * a.path_esize : 0;
        // the dense output is cleared with bulk (TMA) stores from the zeroed shared buffer, a slice per tile:
        // one instruction per 2 KB instead of one 16-byte store per thread (round 1: the staging warps spent
        // a third of their issue slots on those stores)
        const bool zbulk = bulk_zero_ok(pb, pbytes);
        int stage = 0;
        uint32_t phase = 0;
        for (int t = 0; t < ntiles; ++t) {
            if (t >= L.nstages) mbar_wait(&ring.empty[stage], phase ^ 1u);
            float *dst = stages + stage * ring.stage_floats;
            bool async_done = false;
            if (tma) {
                // TMA tensor loads: ONE thread moves the band of the tile as [16 rows x 32 frames] boxes in
                // natural row order (box k = rows lo8 + 16 k ..; lo8 = band start rounded down to the 8-row
                // swizzle atom; rows past the band -- at most 15, possibly the next utterance's -- land in the
                // stage's slack rows and are never read).  Frames beyond T_y read as zero.  No per-thread
                // copy instruction, no LSU miss-queue entry per 16 bytes: the staging warps only clear the
                // output, and one CTA keeps whole tiles in flight (profiles/r2_fast3_phases.txt: per-thread
                // cp.async tops out near 20 GB/s per SM with 128 threads).
                if (tid == 0) {
                    const int y0 = t * kTileY;
                    const int lo8 = max(0, tx + y0 - ty) & ~7;
                    const int hi = min(tx - 1, y0 + kTileY - 1);
                    const int nbox = (hi - lo8) / kTmaBoxRows + 1;
                    mbar_arrive_expect_tx(&ring.full[stage], (uint32_t)nbox * kTmaBoxRows * 128u);
                    for (int k = 0; k < nbox; ++k)
                        tma_load_2d(dst + ((lo8 + kTmaBoxRows * k) << 5), &tmap, y0, b * T_x + lo8 + kTmaBoxRows * k,
                                    &ring.full[stage]);
                }
                async_done = true;
            }
            if constexpr (sizeof(InT) == 4) {
                if (tma) {
                } else if (a.load_mode == 2) {
                    stage_value_tile_async<4>(reinterpret_cast<const float *>(vb), dst, t, tx, ty,
                                              T_y, hw, kHelperWarps, lane);
                    async_done = true;
                } else if (a.load_mode == 4) {
                    stage_value_tile_async<2>(reinterpret_cast<const float *>(vb), dst, t, tx, ty,
                                              T_y, hw, kHelperWarps, lane);
                    async_done = true;
                } else if (a.load_mode == 1) {
                    stage_value_tile_async<1>(reinterpret_cast<const float *>(vb), dst, t, tx, ty,
                                              T_y, hw, kHelperWarps, lane);
                    async_done = true;
                }
            }
            if (tma) {
            } else if (async_done) {
                cp_async_arrive(&ring.full[stage]);
            } else {
                stage_value_tile<InT>(vb, mb, dst, t, tx, ty, T_y, hw, kHelperWarps, lane);
                mbar_arrive(&ring.full[stage]);
            }
            if (++stage == L.nstages) {
                stage = 0;
                phase ^= 1u;
            }
            if (zbulk) zero_fill_bulk_part(pb, pbytes, t, ntiles, zbuf, kFastZeroBytes, htid, nht);
            else zero_fill_part(pb, pbytes, t, ntiles, htid, nht);
        }
        if (ntiles == 0) {
            if (zbulk) zero_fill_bulk_part(pb, pbytes, 0, 1, zbuf, kFastZeroBytes, htid, nht);
            else zero_fill_part(pb, pbytes, 0, 1, htid, nht);
        }
        if (zbulk) {   // the zeros must be in memory before anybody writes a 1-cell
            bulk_commit();
            bulk_wait_all();
        }
    }
    __syncthreads();
    write_path_ones(a.path ? static_cast<char *>(a.path) + (int64_t)b * T_x * T_y * a.path_esize
                           : nullptr,
                    a.durations ? a.durations + (int64_t)b * T_x : nullptr, first, dur, T_x, T_y,
                    a.path_esize, a.one, tid, kFastThreads);
    write_frame_idx(a.frame_idx ? a.frame_idx + (int64_t)b * T_y : nullptr, first, dur, T_x, ty,
                    a.T_y, tid, kFastThreads);
}

// ------------------------------------------------------------------------------------
// fast kernel, long token axis: two DP warps split the tokens (dp_forward2), same staging
// ------------------------------------------------------------------------------------
template <typename InT>
__device__ __forceinline__ void stage_value_tile2(const InT *__restrict__ vb, const float *__restrict__ mb,
                                                  float *stage, int t, int tx, int ty, int64_t T_y, int hw,
                                                  int nhw, int lane, bool async, bool vec16)
{
    // as stage_value_tile / stage_value_tile_async, with the 64-lane row map
    const RowMap rm(tx, 6);
    const int y0 = t * kTileY;
    const int lo = max(0, tx + y0 - ty);
    const int hi = min(tx - 1, y0 + kTileY - 1);
    if (async && vec16) {
        const int c = lane & 7, r = lane >> 3;
        const int left = ty - (y0 + 4 * c);
        const uint32_t bytes = left >= 4 ? 16u : (left > 0 ? 4u * left : 0u);
        const int yo = bytes ? y0 + 4 * c : 0;
        for (int x = lo + 4 * hw + r; x <= hi; x += 4 * nhw) {
            const int row = rm.row(x);
            cp_async16(stage + (row << 5) + ((c ^ (row & 7)) << 2),
                       reinterpret_cast<const float *>(vb) + (int64_t)x * T_y + yo, bytes);
        }
    } else if (async) {
        const int y = y0 + lane;
        const uint32_t bytes = y < ty ? 4u : 0u;
        const float *src = reinterpret_cast<const float *>(vb) + (y < ty ? y : 0);
        for (int x = lo + hw; x <= hi; x += nhw)
            cp_async4(stage + tile_index(rm.row(x), lane), src + (int64_t)x * T_y, bytes);
    } else {
        const int y = y0 + lane;
        const bool in = y < ty;
        for (int x = lo + hw; x <= hi; x += nhw) {
            float v = 0.0f;
            if (in) {
                const int64_t e = (int64_t)x * T_y + y;
                v = load_as_f32(vb + e);
                if (mb) v *= __ldg(mb + e);
            }
            stage[tile_index(rm.row(x), lane)] = v;
        }
    }
}

template <typename InT, int XPLMAX>
__global__ void __launch_bounds__(kFast2Threads) mas_fast2_kernel(const MasArgs a)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const FastLayout &L = a.lay;
    float *stages = reinterpret_cast<float *>(smem + L.off_stages);
    int *first = reinterpret_cast<int *>(smem + L.off_first);
    int *dur = reinterpret_cast<int *>(smem + L.off_dur);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.off_bars);
    uint64_t *edge_full = bars + 12;                                     // [4], after full[<=6] + empty[<=6]
    float *edge = reinterpret_cast<float *>(smem + L.off_bars + 128);    // [4 tiles][32 frames]

    const int b = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T_x = a.T_x;
    const int64_t T_y = a.T_y;
    const int tx = min(max(a.t_x[b], 0), T_x);
    const int ty = min(max(a.t_y[b], 0), a.T_y);
    const bool degenerate = tx > ty && ty >= 1;
    const bool active = tx >= 1 && ty >= 1 && !degenerate;
    const int ntiles = active ? (ty + kTileY - 1) / kTileY : 0;
    constexpr int kHelperWarps = 4;

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
            mbar_init(&ring.full[s], kHelperWarps * 32);
            mbar_init(&ring.empty[s], 2);   // both DP warps
        }
        for (int s = 0; s < 4; ++s) mbar_init(&edge_full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const InT *vb = static_cast<const InT *>(a.value) + (int64_t)b * T_x * T_y;
    const float *mb = a.cell_mask ? a.cell_mask + (int64_t)b * T_x * T_y : nullptr;

    if (warp >= kHelperWarps) {
        // ---------------- the two DP warps ----------------
        const int w = warp - kHelperWarps;
        if (w == 0)
            for (int x = lane; x < T_x; x += 32) dur[x] = 0;
        __syncwarp();
        float score = 0.0f;
        int owns = 0;
        if (active) {
            score = prior_forward2_dispatch<XPLMAX>(ring, bits, L.xrows, tx, ty, lane, w, 0, edge, edge_full,
                                                    &owns, nullptr);
            if (owns && lane == 0 && a.score) a.score[b] = score;
            __threadfence_block();
            asm volatile("bar.sync 1, 64;" ::: "memory");   // both halves of the direction bits are written
            if (w == 0 && lane == 0) backtrack_bits(bits, L.xrows, tx, ty, first, dur, L.bits_in_smem != 0, 6);
        } else if (w == 0) {
            if (degenerate) {
                if (lane == 0) {
                    auto val = [&](int x, int y) {
                        const int64_t e = (int64_t)x * T_y + y;
                        float v = load_as_f32(vb + e);
                        if (mb) v *= mb[e];
                        return v;
                    };
                    backtrack_degenerate(val, tx, ty, first, dur);
                    score = val(tx - 1, ty - 1);
                }
                score = __shfl_sync(kFull, score, 0);
            }
            if (lane == 0 && a.score) a.score[b] = score;
        }
    } else {
        // ---------------- staging warps ----------------
        const int hw = warp;
        constexpr int nht = kHelperWarps * 32;
        char *pb = a.path ? static_cast<char *>(a.path) + (int64_t)b * T_x * T_y * a.path_esize : nullptr;
        const int64_t pbytes = a.path ? (int64_t)T_x * T_y * a.path_esize : 0;
        const bool async = sizeof(InT) == 4 && a.load_mode != 0;
        int stage = 0;
        uint32_t phase = 0;
        for (int t = 0; t < ntiles; ++t) {
            if (t >= L.nstages) mbar_wait(&ring.empty[stage], phase ^ 1u);
            float *dst = stages + stage * ring.stage_floats;
            if (async && a.load_mode == 4)
                stage_value_tile_async<2, 6>(reinterpret_cast<const float *>(vb), dst, t, tx, ty, T_y, hw, kHelperWarps, lane);
            else
                stage_value_tile2<InT>(vb, mb, dst, t, tx, ty, T_y, hw, kHelperWarps, lane, async, a.load_mode == 2);
            if (async) cp_async_arrive(&ring.full[stage]);
            else mbar_arrive(&ring.full[stage]);
            if (++stage == L.nstages) {
                stage = 0;
                phase ^= 1u;
            }
            zero_fill_part(pb, pbytes, t, ntiles, tid, nht);
        }
        if (ntiles == 0) zero_fill_part(pb, pbytes, 0, 1, tid, nht);
    }
    __syncthreads();
    write_path_ones(a.path ? static_cast<char *>(a.path) + (int64_t)b * T_x * T_y * a.path_esize : nullptr,
                    a.durations ? a.durations + (int64_t)b * T_x : nullptr, first, dur, T_x, T_y, a.path_esize,
                    a.one, tid, kFast2Threads);
    write_frame_idx(a.frame_idx ? a.frame_idx + (int64_t)b * T_y : nullptr, first, dur, T_x, ty, a.T_y, tid,
                    kFast2Threads);
}

template <typename InT>
static cudaError_t launch_fast2_typed(const MasArgs &a, cudaStream_t st)
{
    const int xplmax = (a.T_x + 63) / 64;
    void (*k)(const MasArgs) = (xplmax <= 6) ? mas_fast2_kernel<InT, 6> : mas_fast2_kernel<InT, 8>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.lay.total);
    if (e != cudaSuccess) return e;
    k<<<a.B, kFast2Threads, a.lay.total, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

template <typename InT>
static cudaError_t launch_fast_typed(const MasArgs &a, cudaStream_t st, const TensorMap *tmap)
{
    const int xplmax = (a.T_x + 31) / 32;
    void (*k)(const MasArgs, const TensorMap) = nullptr;
    if constexpr (sizeof(InT) == 4) {  // fp32: register budget sized to the batch's T_x bucket
        if (xplmax <= 2) k = mas_fast_kernel<InT, 2>;
        else if (xplmax <= 4) k = mas_fast_kernel<InT, 4>;
        else if (xplmax <= 6) k = mas_fast_kernel<InT, 6>;
        else if (xplmax <= 8) k = mas_fast_kernel<InT, 8>;
        else if (xplmax <= 12) k = mas_fast_kernel<InT, 12>;
        else k = mas_fast_kernel<InT, 16>;
    } else {  // fp16 / bf16 / fp64 inputs: two buckets keep the library small
        if (xplmax <= 8) k = mas_fast_kernel<InT, 8>;
        else k = mas_fast_kernel<InT, 16>;
    }
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)a.lay.total);
    if (e != cudaSuccess) return e;
    static const TensorMap no_map{};
    k<<<a.B, kFastThreads, a.lay.total, st>>>(a, tmap ? *tmap : no_map);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_fast(const MasArgs &a, int value_dtype, cudaStream_t st, const TensorMap *tmap)
{
    if (a.dp_warps == 2) {
        switch (value_dtype) {
        case MAS_F32: return launch_fast2_typed<float>(a, st);
        case MAS_F16: return launch_fast2_typed<__half>(a, st);
        case MAS_BF16: return launch_fast2_typed<__nv_bfloat16>(a, st);
        case MAS_F64: return launch_fast2_typed<double>(a, st);
        default: return cudaErrorInvalidValue;
        }
    }
    switch (value_dtype) {
    case MAS_F32: return launch_fast_typed<float>(a, st, tmap);
    case MAS_F16: return launch_fast_typed<__half>(a, st, tmap);
    case MAS_BF16: return launch_fast_typed<__nv_bfloat16>(a, st, tmap);
    case MAS_F64: return launch_fast_typed<double>(a, st, tmap);
    default: return cudaErrorInvalidValue;
    }
}

// ------------------------------------------------------------------------------------
// general kernel (any T_x): block-wide sweep, bits [frame][token word] in the workspace
// ------------------------------------------------------------------------------------
template <typename InT>
__global__ void __launch_bounds__(kGeneralThreads) mas_general_kernel(const MasArgs a)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int T_x = a.T_x;
    const int64_t T_y = a.T_y;
    float *Va = reinterpret_cast<float *>(smem);
    float *Vb = Va + T_x;
    int *first = reinterpret_cast<int *>(Vb + T_x);
    int *dur = first + T_x;

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int tx = min(max(a.t_x[b], 0), T_x);
    const int ty = min(max(a.t_y[b], 0), a.T_y);
    const bool degenerate = tx > ty && ty >= 1;
    const bool active = tx >= 1 && ty >= 1 && !degenerate;
    const int xw = (T_x + 31) >> 5;
    uint32_t *bits = a.bits_ws + (size_t)b * (size_t)a.lay.nch * 32 * xw;
    const InT *vb = static_cast<const InT *>(a.value) + (int64_t)b * T_x * T_y;
    const float *mb = a.cell_mask ? a.cell_mask + (int64_t)b * T_x * T_y : nullptr;
    char *pb = a.path ? static_cast<char *>(a.path) + (int64_t)b * T_x * T_y * a.path_esize : nullptr;

    for (int x = tid; x < T_x; x += kGeneralThreads) {
        Va[x] = kNeg;
        Vb[x] = kNeg;
        dur[x] = 0;
    }
    zero_fill_part(pb, a.path ? (int64_t)T_x * T_y * a.path_esize : 0, 0, 1, tid, kGeneralThreads);
    __syncthreads();

    float *Vp = Va, *Vc = Vb;
    if (active) {
        for (int y = 0; y < ty; ++y) {
            const int lo = max(0, tx + y - ty), hi = min(tx - 1, y);
            const int xb0 = (lo >> 5) << 5;
            for (int xb = xb0 + (tid & ~31); xb <= hi; xb += kGeneralThreads) {
                const int x = xb + lane;
                bool bit = false;
                if (x >= lo && x <= hi) {
                    const float vc = (x == y) ? kNeg : Vp[x];
                    const float vp = (x == 0) ? ((y == 0) ? 0.0f : kNeg) : Vp[x - 1];
                    const bool take_prev = vp > vc;
                    const int64_t e = (int64_t)x * T_y + y;
                    float v = load_as_f32(vb + e);
                    if (mb) v *= mb[e];
                    Vc[x] = __fadd_rn(take_prev ? vp : vc, v);
                    bit = (x != 0) && (x == y || take_prev);
                }
                const uint32_t w = __ballot_sync(kFull, bit);
                if (lane == 0) bits[(size_t)y * xw + (xb >> 5)] = w;
            }
            __syncthreads();
            float *tmp = Vp;
            Vp = Vc;
            Vc = tmp;
        }
        if (tid == 0) {
            if (a.score) a.score[b] = Vp[tx - 1];
            int idx = tx - 1, top = ty - 1;
            for (int y = ty - 1; y >= 0; --y) {
                const uint32_t w = bits[(size_t)y * xw + (idx >> 5)];
                if ((w >> (idx & 31)) & 1u) {
                    first[idx] = y;
                    dur[idx] = top - y + 1;
                    --idx;
                    top = y - 1;
                }
            }
            if (top >= 0) {
                first[idx] = 0;
                dur[idx] = top + 1;
            }
        }
    } else if (tid == 0) {
        float s = 0.0f;
        if (degenerate) {
            auto val = [&](int x, int y) {
                const int64_t e = (int64_t)x * T_y + y;
                float v = load_as_f32(vb + e);
                if (mb) v *= mb[e];
                return v;
            };
            backtrack_degenerate(val, tx, ty, first, dur);
            s = val(tx - 1, ty - 1);
        }
        if (a.score) a.score[b] = s;
    }
    __syncthreads();
    write_path_ones(pb, a.durations ? a.durations + (int64_t)b * T_x : nullptr, first, dur, T_x, T_y,
                    a.path_esize, a.one, tid, kGeneralThreads);
    write_frame_idx(a.frame_idx ? a.frame_idx + (int64_t)b * T_y : nullptr, first, dur, T_x, ty,
                    a.T_y, tid, kGeneralThreads);
}

cudaError_t launch_general(const MasArgs &a, int value_dtype, cudaStream_t st)
{
    const size_t smem = (size_t)a.T_x * 16;
    void (*k)(const MasArgs) = nullptr;
    switch (value_dtype) {
    case MAS_F32: k = mas_general_kernel<float>; break;
    case MAS_F16: k = mas_general_kernel<__half>; break;
    case MAS_BF16: k = mas_general_kernel<__nv_bfloat16>; break;
    case MAS_F64: k = mas_general_kernel<double>; break;
    default: return cudaErrorInvalidValue;
    }
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k<<<a.B, kGeneralThreads, smem, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// generate_path (src/model/utils.py:26-43): durations -> dense path
// ------------------------------------------------------------------------------------
// grid (B, nsplit): every CTA recomputes the (tiny) cumulative sum, then owns a contiguous
// slab of token rows: zero-fills it with 16-byte stores and writes the runs of ones.
__global__ void __launch_bounds__(256) generate_path_kernel(const void *dur_in, int dur_dtype,
                                                            const int32_t *t_x, const int32_t *t_y,
                                                            void *path, int esize,
                                                            unsigned long long one, int T_x, int T_y)
{
    extern __shared__ __align__(16) unsigned char smem[];
    int *start = reinterpret_cast<int *>(smem);  // [T_x + 1] first frame of every token
    const int b = blockIdx.x, tid = threadIdx.x;
    const int tx = t_x ? min(max(t_x[b], 0), T_x) : T_x;
    const int ty = t_y ? min(max(t_y[b], 0), T_y) : T_y;
    if (tid < 32) {
        // cumsum by one warp, sequential over 32-token groups; fp32 durations are accumulated left to right in
        // DOUBLE and every partial sum rounded to fp32 -- what torch.cumsum does on the host (ATen's CPU
        // accumulation type of float is double; captured from the reference's forward() with a fractional
        // length_scale, tests/golden/inference_arttts.npz) -- then y < cum  <=>  y < ceil(cum)
        const int lane = tid;
        if (dur_dtype == MAS_I32) {
            const int32_t *d = static_cast<const int32_t *>(dur_in) + (int64_t)b * T_x;
            int carry = 0;
            for (int x0 = 0; x0 < T_x; x0 += 32) {
                const int x = x0 + lane;
                int v = (x < T_x) ? d[x] : 0;
                for (int o = 1; o < 32; o <<= 1) {
                    const int n = __shfl_up_sync(kFull, v, o);
                    if (lane >= o) v += n;
                }
                if (x < T_x) start[x + 1] = min(max(carry + v, 0), T_y);
                carry += __shfl_sync(kFull, v, 31);
            }
        } else if (lane == 0) {
            const float *d = static_cast<const float *>(dur_in) + (int64_t)b * T_x;
            double acc = 0.0;
            for (int x = 0; x < T_x; ++x) {
                acc += (double)d[x];
                const float c = ceilf((float)acc);
                start[x + 1] = (c <= 0.0f) ? 0 : (c >= (float)T_y ? T_y : (int)c);
            }
        }
        if (lane == 0) start[0] = 0;
    }
    __syncthreads();
    const int nsplit = gridDim.y;
    const int rows_per = (T_x + nsplit - 1) / nsplit;
    const int x0 = blockIdx.y * rows_per, x1 = min(T_x, x0 + rows_per);
    if (x0 >= x1) return;
    char *pb = static_cast<char *>(path) + ((int64_t)b * T_x + x0) * (int64_t)T_y * esize;
    zero_fill_part(pb, (int64_t)(x1 - x0) * T_y * esize, 0, 1, tid, 256);
    __syncthreads();
    for (int x = x0 + tid; x < x1; x += 256) {
        if (x >= tx) continue;
        const int f = start[x], l = min(start[x + 1], ty);
        for (int y = f; y < l; ++y) st_one(pb, (int64_t)(x - x0) * T_y + y, esize, one);
    }
}

cudaError_t launch_generate_path(const void *dur, int dur_dtype, const int32_t *t_x,
                                 const int32_t *t_y, void *path, int esize, unsigned long long one,
                                 int B, int T_x, int T_y, cudaStream_t st)
{
    int nsplit = 1;
    while (B * nsplit < 296 && nsplit * 8 <= T_x) nsplit *= 2;
    const size_t smem = (size_t)(T_x + 1) * sizeof(int);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(generate_path_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    generate_path_kernel<<<dim3(B, nsplit), 256, smem, st>>>(dur, dur_dtype, t_x, t_y, path, esize,
                                                             one, T_x, T_y);
    count_launch();
    return cudaGetLastError();
}

}  // namespace mas
