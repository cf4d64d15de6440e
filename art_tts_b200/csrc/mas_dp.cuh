// mas_dp.cuh -- the single-warp MAS recurrence, direction-bit packing and the backtrack.
//
// Reference semantics being reproduced (paths relative to /root/reference/):
//   forward   src/model/monotonic_align/core.pyx:17-30
//   backtrack src/model/monotonic_align/core.pyx:15,32-35
//
// Formulation (validated against the reference kernel on CPU in
// oracle.maximum_path_rowsweep):  one fp32 score per token lives in a register and is
// updated once per frame,
//     V[x] <- ((V[x-1] > V[x]) ? V[x-1] : V[x]) + value[x,y],     V[x] == -1e9 while x > y,
// and the only thing that outlives a frame is the backtrack predicate
//     d[x,y] = (x != 0) && (x == y || V[x,y-1] < V[x-1,y-1]).
//
// Token ownership: lane l owns the XPL CONSECUTIVE tokens x = l*XPL + j (XPL = ceil(t_x/32),
// per utterance).  The x-1 neighbour is then a register for all but the lane's first token,
// and that one needs the left lane's LAST token of the previous frame -- a value produced
// a whole frame earlier, so its shuffle is issued right after it is computed and lands
// while the lane's other tokens are being updated: the recurrence is issue-bound, not
// shuffle-latency-bound (measured 120 + 28*XPL cycles/frame with the strided ownership
// this replaced).  Staged tiles keep token x in physical row (x % XPL)*32 + x / XPL so that
// lane l still reads rows l, 32+l, 64+l, ... : conflict-free LDS.128 under the 128-byte
// XOR swizzle.  A 32-frame tile of d packs into one 32-bit word per token without any
// cross-lane traffic (bit s of word [chunk][row] is frame 32*chunk+s).
#pragma once

#include "mas_common.cuh"

namespace mas {

// Per-utterance token -> physical row mapping (see above).  `inv` = ceil(65536 / XPL) makes
// x / XPL an exact multiply-shift for x < 512, XPL <= 16.
// `lsh` = log2 of the lanes that share an utterance: 5 for the single-warp recurrence, 6 when two
// DP warps split the token axis (dp_forward2): lane L of 2^lsh owns tokens L*xpl .. L*xpl+xpl-1.
struct RowMap {
    int xpl;
    uint32_t inv;
    int lsh;
    __device__ __forceinline__ explicit RowMap(int tx, int lane_shift = 5)
    {
        lsh = lane_shift;
        xpl = max(1, (tx + (1 << lsh) - 1) >> lsh);
        inv = (65536u + xpl - 1) / xpl;
    }
    __device__ __forceinline__ int row(int x) const
    {
        const int q = (int)((uint32_t)x * inv >> 16);
        return ((x - q * xpl) << lsh) + q;
    }
};

template <int K>
__device__ __forceinline__ float f4_get(const float4 &v)
{
    if constexpr (K == 0) return v.x;
    if constexpr (K == 1) return v.y;
    if constexpr (K == 2) return v.z;
    return v.w;
}

// direction bit of one cell: acc |= bit when up > cur (the comparison of core.pyx:30), as FSETP +
// a predicated integer op that sit beside the value chain, not in it
__device__ __forceinline__ void dir_bit(float up, float cur, uint32_t &acc, uint32_t bit)
{
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(acc) : "f"(up), "f"(cur), "r"(bit));
}

// One frame.  `v[j]` is value[x_j, y]; `bit` = 1 << (y & 31); `left` is the left lane's last
// token at frame y-1 (lane 0: the x == 0 boundary of core.pyx:23-27) and is replaced by the
// value for frame y.  Tokens are updated from the lane's last to its first so that V[j-1]
// is still the previous frame's value when token j reads it.
template <int XPL, bool DIAG, bool FMAX = false>   // FMAX: see dp_step2
__device__ __forceinline__ void dp_step(float (&V)[XPL], uint32_t (&acc)[XPL],
                                        const float (&v)[XPL], float &left, int lane, int x0,
                                        int y, uint32_t bit)
{
    float nxt = 0.0f;
#pragma unroll
    for (int j = XPL - 1; j >= 0; --j) {
        const float up = (j == 0) ? left : V[j > 0 ? j - 1 : 0];   // V[x-1, y-1]
        float m;
        if (FMAX) {
            dir_bit(up, V[j], acc[j], bit);
            m = fmaxf(up, V[j]);
        } else {
            const bool take_prev = up > V[j];                      // core.pyx:30 max()
            m = take_prev ? up : V[j];
            if (take_prev) acc[j] |= bit;
        }
        float nv = __fadd_rn(m, v[j]);
        if (DIAG) nv = (x0 + j <= y) ? nv : kNeg;                  // x > y: not reachable yet
        V[j] = nv;
        if (j == XPL - 1) nxt = __shfl_up_sync(kFull, nv, 1);      // in flight during the rest
    }
    left = (lane == 0) ? kNeg : nxt;  // token 0: v_prev = -1e9 after the first frame
}

// One staged tile of up to 32 frames (frames y0 .. y0+nsteps-1, y0 % 32 == 0).
// `stage` is the swizzled [row][32 frames] tile in shared memory (see tile_index()).
// NAT: the tile keeps token x in row x (what a TMA box writes, SWIZZLE_128B: chunk ^= row & 7) instead of
// the permuted row (x % XPL) * 32 + x / XPL of the cp.async / STS staging.
template <int XPL, bool DIAG, bool FULL, bool FMAX = false, bool NAT = false>
__device__ __forceinline__ void dp_tile(float (&V)[XPL], uint32_t (&acc)[XPL], float &left,
                                        const float *__restrict__ stage, int lane, int x0, int y0,
                                        int nsteps)
{
    // explicit shared-space address: `stage` arrives through a struct, and a generic LD would
    // otherwise be emitted for the hottest load of the kernel
    const uint32_t rowbase = smem_u32(stage + ((NAT ? x0 : lane) << 5));
    const int sw = lane & 7;
#pragma unroll 1
    for (int g = 0; g < 8; ++g) {
        const int s0 = g << 2;
        if (!FULL && s0 >= nsteps) break;
        float4 vv[XPL];
#pragma unroll
        for (int j = 0; j < XPL; ++j)
            vv[j] = NAT ? lds128(rowbase + (j << 7) + ((g ^ ((x0 + j) & 7)) << 4))
                        : lds128(rowbase + (j << 12) + ((g ^ sw) << 4));
        float v[XPL];
#pragma unroll
        for (int j = 0; j < XPL; ++j) v[j] = f4_get<0>(vv[j]);
        dp_step<XPL, DIAG, FMAX>(V, acc, v, left, lane, x0, y0 + s0, 1u << s0);
        if (FULL || s0 + 1 < nsteps) {
#pragma unroll
            for (int j = 0; j < XPL; ++j) v[j] = f4_get<1>(vv[j]);
            dp_step<XPL, DIAG, FMAX>(V, acc, v, left, lane, x0, y0 + s0 + 1, 2u << s0);
        }
        if (FULL || s0 + 2 < nsteps) {
#pragma unroll
            for (int j = 0; j < XPL; ++j) v[j] = f4_get<2>(vv[j]);
            dp_step<XPL, DIAG, FMAX>(V, acc, v, left, lane, x0, y0 + s0 + 2, 4u << s0);
        }
        if (FULL || s0 + 3 < nsteps) {
#pragma unroll
            for (int j = 0; j < XPL; ++j) v[j] = f4_get<3>(vv[j]);
            dp_step<XPL, DIAG, FMAX>(V, acc, v, left, lane, x0, y0 + s0 + 3, 8u << s0);
        }
    }
}

// Producer/consumer ring shared by the drop-in kernel (tiles = value) and the fused kernel
// (tiles = log-prior computed on the fly).
struct TileRing {
    float *stages;       // nstages x (xrows*32) floats, 1024-byte aligned
    uint64_t *full;      // [nstages] producers -> DP warp
    uint64_t *empty;     // [nstages] DP warp -> producers
    int nstages;
    int stage_floats;    // xrows * 32
};

// Forward pass of one utterance by ONE warp.  Consumes ceil(t_y/32) tiles from the ring,
// writes the direction words to bits[chunk*xrows + row] (shared or global memory) and
// returns V[t_x-1, t_y-1].  Requires 1 <= t_x <= t_y and XPL == ceil(t_x/32).
// `g0` = index (in the ring's lifetime) of this utterance's first tile: a persistent CTA keeps
// one ring and its barrier phases running across utterances.
template <int XPL, bool FMAX = false, bool NAT = false>
__device__ __noinline__ float dp_forward(const TileRing ring, uint32_t *bits, int xrows, int tx,
                                         int ty, int lane, int g0 = 0, long long *wait_acc = nullptr)
{
    float V[XPL];
    uint32_t acc[XPL];
#pragma unroll
    for (int j = 0; j < XPL; ++j) {
        V[j] = kNeg;
        acc[j] = 0u;
    }
    const int x0 = lane * XPL;                  // this lane's first token
    float left = (lane == 0) ? 0.0f : kNeg;     // frame 0: v_prev(x=0) = 0, everything else -1e9
    const int ntiles = (ty + kTileY - 1) / kTileY;
    int stage = g0 % ring.nstages;
    uint32_t phase = (uint32_t)(g0 / ring.nstages) & 1u;
    for (int t = 0; t < ntiles; ++t) {
        if (wait_acc) {  // profiling aid: cycles this warp spends starved of tiles
            const long long t0 = clock64();
            mbar_wait(&ring.full[stage], phase);
            *wait_acc += clock64() - t0;
        } else {
            mbar_wait(&ring.full[stage], phase);
        }
        const float *tile = ring.stages + stage * ring.stage_floats;
        const int y0 = t * kTileY;
        const int nsteps = min(kTileY, ty - y0);
        const bool diag = y0 < tx;  // some token x > y still exists in this tile
        if (nsteps == kTileY) {
            if (diag) dp_tile<XPL, true, true, FMAX, NAT>(V, acc, left, tile, lane, x0, y0, nsteps);
            else dp_tile<XPL, false, true, FMAX, NAT>(V, acc, left, tile, lane, x0, y0, nsteps);
        } else {
            dp_tile<XPL, true, false, FMAX, NAT>(V, acc, left, tile, lane, x0, y0, nsteps);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ring.empty[stage]);
        if (++stage == ring.nstages) {
            stage = 0;
            phase ^= 1u;
        }
        // x == y always steps down (core.pyx:34 `index == y`), token 0 never does.
        if (diag) {
#pragma unroll
            for (int j = 0; j < XPL; ++j)
                if (((x0 + j) >> 5) == t) acc[j] |= 1u << ((x0 + j) & 31);
        }
        if (lane == 0) acc[0] = 0u;
        uint32_t *dst = bits + (size_t)t * xrows + lane;
        MAS_CHECK(((XPL - 1) << 5) + lane < xrows && t < ntiles);
#pragma unroll
        for (int j = 0; j < XPL; ++j) {
            dst[j << 5] = acc[j];
            acc[j] = 0u;
        }
    }
    // total alignment score: token tx-1 = lane (tx-1)/XPL, slot (tx-1)%XPL
    const int ql = (tx - 1) / XPL, qj = (tx - 1) - ql * XPL;
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < XPL; ++j)
        if (j == qj) s = V[j];
    return __shfl_sync(kFull, s, ql);
}

// Backtrack over the packed direction words (core.pyx:32-35), executed by one lane.
// Instead of one step per frame it jumps from token boundary to token boundary: inside a
// 32-frame word the next decrement is the highest set bit at or below the current frame.
// The word of the NEXT token (same chunk) is prefetched every iteration, so the common
// transition (next token, same chunk) costs a handful of dependent ALU ops, not a load.
// Records, per visited token, its first frame and its duration.
template <bool SMEM>
__device__ __forceinline__ uint32_t bt_word(const uint32_t *bits, uint32_t sbits, int xrows, int c,
                                            int row)
{
    if constexpr (SMEM) {
        uint32_t w;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(sbits + 4u * (uint32_t)(c * xrows + row)));
        return w;
    } else {
        return bits[(size_t)c * xrows + row];
    }
}

template <bool SMEM>
__device__ __forceinline__ void backtrack_bits_impl(const uint32_t *bits, int xrows, int tx, int ty,
                                                    int *first, int *dur, int lsh = 5)
{
    const RowMap rm(tx, lsh);
    const int xpl = rm.xpl;
    const uint32_t sbits = SMEM ? smem_u32(bits) : 0u;
    int idx = tx - 1, y = ty - 1, top = ty - 1;
    int q = idx / xpl, j = idx - q * xpl;          // token idx = lane q, slot j -> row (j<<5)+q
    int row = (j << lsh) + q;
    // row of token idx-1
    auto step_row = [&](int &jj, int &qq) {
        if (jj > 0) --jj;
        else { jj = xpl - 1; --qq; }
        return (jj << lsh) + qq;
    };
    int jn = j, qn = q;
    int rown = idx > 0 ? step_row(jn, qn) : 0;
    int c = y >> 5;
    uint32_t w = bt_word<SMEM>(bits, sbits, xrows, c, row);
    uint32_t wn = idx > 0 ? bt_word<SMEM>(bits, sbits, xrows, c, rown) : 0u;
    while (y >= 0) {
        const int s = y & 31;
        const uint32_t m = (idx != 0 ? w : 0u) & (0xffffffffu >> (31 - s));
        if (m == 0u) {  // stays on this token down to the chunk start
            y = (c << 5) - 1;
            if (y < 0) break;
            --c;
            w = bt_word<SMEM>(bits, sbits, xrows, c, row);
            wn = idx > 0 ? bt_word<SMEM>(bits, sbits, xrows, c, rown) : 0u;
        } else {
            const int p = 31 - __clz(m);
            const int ys = (c << 5) + p;
            MAS_CHECK(idx > 0 && idx < tx && c >= 0 && ys <= top && row >= 0 && row < xrows);
            first[idx] = ys;
            dur[idx] = top - ys + 1;
            --idx;
            y = ys - 1;
            top = y;
            row = rown;
            w = wn;
            if (idx > 0) rown = step_row(jn, qn);
            if (p == 0) {  // crossed into the previous chunk
                if (y < 0) break;
                --c;
                w = bt_word<SMEM>(bits, sbits, xrows, c, row);
            }
            wn = idx > 0 ? bt_word<SMEM>(bits, sbits, xrows, c, rown) : 0u;
        }
    }
    if (top >= 0) {
        first[idx] = 0;
        dur[idx] = top + 1;
    }
}

__device__ __forceinline__ void backtrack_bits(const uint32_t *bits, int xrows, int tx, int ty,
                                               int *first, int *dur, bool bits_in_smem, int lsh = 5)
{
    if (bits_in_smem) backtrack_bits_impl<true>(bits, xrows, tx, ty, first, dur, lsh);
    else backtrack_bits_impl<false>(bits, xrows, tx, ty, first, dur, lsh);
}

// The reference's degenerate case t_x > t_y: the band of core.pyx:18 is empty, the
// forward pass changes nothing and the backtrack compares RAW (masked) values.  Rare, so
// one lane simply evaluates the two cells it needs per frame (`val(x, y)`).  index never
// reaches y or 0 here.
template <typename ValFn>
__device__ __forceinline__ void backtrack_degenerate(ValFn val, int tx, int ty, int *first, int *dur)
{
    int idx = tx - 1, top = ty - 1;
    for (int y = ty - 1; y > 0 && idx > 0; --y) {
        const float a = val(idx, y - 1), b = val(idx - 1, y - 1);
        if (a < b) {
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

// After the block-wide barrier: every thread publishes durations and the 1-cells of the
// (already zero-filled) dense path.
__device__ __forceinline__ void write_path_ones(void *path_b, int32_t *dur_out_b, const int *first,
                                                const int *dur, int T_x, int64_t T_y, int esize,
                                                uint64_t one, int tid, int nthr)
{
    for (int x = tid; x < T_x; x += nthr) {
        const int d = dur[x];
        MAS_CHECK(d >= 0 && d <= T_y);
        MAS_CHECK(d == 0 || (first[x] >= 0 && (int64_t)first[x] + d <= T_y));
        if (dur_out_b) dur_out_b[x] = d;
        if (path_b && d > 0) {
            const int64_t e0 = (int64_t)x * T_y + first[x];
            for (int k = 0; k < d; ++k) st_one(path_b, e0 + k, esize, one);
        }
    }
}

// token index of every frame (-1 on padding) from the per-token runs
__device__ __forceinline__ void write_frame_idx(int32_t *fi, const int *first, const int *dur,
                                                int T_x, int ty, int T_y, int tid, int nthr)
{
    if (!fi) return;
    for (int y = ty + tid; y < T_y; y += nthr) fi[y] = -1;
    for (int x = tid; x < T_x; x += nthr) {
        const int d = dur[x], f0 = first[x];
        MAS_CHECK(d == 0 || (f0 >= 0 && f0 + d <= ty));
        for (int k = 0; k < d; ++k) fi[f0 + k] = x;
    }
}

// ------------------------------------------------------------------------------------
// Two DP warps per utterance.  The 64 lanes of warps w = 0, 1 own consecutive token ranges
// (global lane L = 32w + lane owns tokens L*XPL .. L*XPL+XPL-1, XPL = ceil(t_x/64)), so each warp
// runs the recurrence of dp_step on half the tokens.  The one value that crosses the warp
// boundary -- warp 0's last token of frame y, which warp 1's first token needs at frame y+1 --
// travels through a small ring in shared memory (`edge`, 4 tiles x 32 frames) and warp 1 simply
// runs at least one tile behind warp 0 (mbarrier `edge_full` per tile): no per-frame handshake.
// Warp 0 cannot lap warp 1 by more than the tile ring's depth (<= 3), so the 4-tile edge ring
// needs no back-pressure of its own.  Tiles use the row map RowMap(tx, 6).
// ------------------------------------------------------------------------------------
// FMAX = the formulation of the fused kernels: the value chain is FMNMX + FADD (10 cycles per
// token instead of 14 for FSETP + FSEL + FADD, profiles/microbench/dp_chain.cu) and the direction
// bit comes from the same comparison beside the chain.  max(up, cur) is the reference's
// `up > cur ? up : cur` on finite values (a NaN or a +0/-0 tie could differ), so the drop-in
// kernel, whose contract is bit-exactness on any input, keeps the select.
#ifdef MAS_TC_TRACE
// per-tile event timestamps of ONE CTA (profiles/tc_trace.py): tr[g * 16 + event], set by mas_prior_tc_kernel
static __device__ long long *g_tc_trace;
#define MAS_TRACE(ev, g)                                                                       \
    do {                                                                                       \
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && g_tc_trace && (g) < 1500) g_tc_trace[(g) * 16 + (ev)] = clock64(); \
    } while (0)
#else
#define MAS_TRACE(ev, g) ((void)0)
#endif

template <int XPL, bool DIAG, bool FMAX = false>
__device__ __forceinline__ float dp_step2(float (&V)[XPL], uint32_t (&acc)[XPL], const float (&v)[XPL],
                                          float &left, int lane, int x0, int y, uint32_t bit,
                                          float lane0_left)
{
    float nxt = 0.0f, last = 0.0f;
#pragma unroll
    for (int j = XPL - 1; j >= 0; --j) {
        const float up = (j == 0) ? left : V[j > 0 ? j - 1 : 0];   // V[x-1, y-1]
        float m;
        if (FMAX) {
            dir_bit(up, V[j], acc[j], bit);
            m = fmaxf(up, V[j]);
        } else {
            const bool take_prev = up > V[j];                      // core.pyx:30 max()
            m = take_prev ? up : V[j];
            if (take_prev) acc[j] |= bit;
        }
        float nv = __fadd_rn(m, v[j]);
        if (DIAG) nv = (x0 + j <= y) ? nv : kNeg;                  // x > y: not reachable yet
        V[j] = nv;
        if (j == XPL - 1) {
            last = nv;
            nxt = __shfl_up_sync(kFull, nv, 1);                    // in flight during the rest
        }
    }
    left = (lane == 0) ? lane0_left : nxt;
    return last;
}

template <int XPL, bool DIAG, bool FULL, bool FMAX = false>
__device__ __forceinline__ void dp_tile2(float (&V)[XPL], uint32_t (&acc)[XPL], float &left,
                                         const float *__restrict__ stage, float *__restrict__ edge_tile,
                                         int w, int lane, int x0, int y0, int nsteps)
{
    const uint32_t rowbase = smem_u32(stage + ((32 * w + lane) << 5));
    const uint32_t edgebase = smem_u32(edge_tile);
    const int sw = lane & 7;
    // 8 frames per trip in the drop-in kernel (cfg1-3: 4-6 % faster); 4 in the fused kernels, whose fourteen warps
    // share the instruction cache (the longer body costs them 3 %: 0.3306 vs 0.3196 ms)
    constexpr int kTrip = FMAX ? 1 : 2;
#pragma unroll kTrip
    for (int g = 0; g < 8; ++g) {
        const int s0 = g << 2;
        if (!FULL && s0 >= nsteps) break;
        float4 vv[XPL];
#pragma unroll
        for (int j = 0; j < XPL; ++j) vv[j] = lds128(rowbase + (j << 13) + ((g ^ sw) << 4));   // rows j*64 + L
        // warp 1: warp 0's last token at these four frames (already published: it is a tile ahead)
        const float4 ev = (w == 1) ? lds128(edgebase + (g << 4)) : make_float4(kNeg, kNeg, kNeg, kNeg);
        float4 out;
        float v[XPL];
#pragma unroll
        for (int j = 0; j < XPL; ++j) v[j] = f4_get<0>(vv[j]);
        out.x = dp_step2<XPL, DIAG, FMAX>(V, acc, v, left, lane, x0, y0 + s0, 1u << s0, ev.x);
        out.y = out.z = out.w = 0.0f;
        if (FULL || s0 + 1 < nsteps) {
#pragma unroll
            for (int j = 0; j < XPL; ++j) v[j] = f4_get<1>(vv[j]);
            out.y = dp_step2<XPL, DIAG, FMAX>(V, acc, v, left, lane, x0, y0 + s0 + 1, 2u << s0, ev.y);
        }
        if (FULL || s0 + 2 < nsteps) {
#pragma unroll
            for (int j = 0; j < XPL; ++j) v[j] = f4_get<2>(vv[j]);
            out.z = dp_step2<XPL, DIAG, FMAX>(V, acc, v, left, lane, x0, y0 + s0 + 2, 4u << s0, ev.z);
        }
        if (FULL || s0 + 3 < nsteps) {
#pragma unroll
            for (int j = 0; j < XPL; ++j) v[j] = f4_get<3>(vv[j]);
            out.w = dp_step2<XPL, DIAG, FMAX>(V, acc, v, left, lane, x0, y0 + s0 + 3, 8u << s0, ev.w);
        }
        if (w == 0 && lane == 31)   // publish my last token (token 32*XPL - 1) for warp 1
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(edgebase + (g << 4)), "f"(out.x),
                         "f"(out.y), "f"(out.z), "f"(out.w)
                         : "memory");
    }
}

// Forward pass of one utterance by DP warp `w` of two.  `edge` = 4 x 32 floats, `edge_full` = 4
// mbarriers (count 1); ring.empty barriers must expect TWO arrivals per tile.  Returns
// V[t_x-1, t_y-1] in every lane of the warp that owns token t_x-1 (`*owns` = 1), else *owns = 0.
template <int XPL, bool FMAX = false>
__device__ __noinline__ float dp_forward2(const TileRing ring, uint32_t *bits, int xrows, int tx, int ty,
                                          int lane, int w, int g0, float *edge, uint64_t *edge_full,
                                          int *owns, long long *wait_acc = nullptr)
{
    float V[XPL];
    uint32_t acc[XPL];
#pragma unroll
    for (int j = 0; j < XPL; ++j) {
        V[j] = kNeg;
        acc[j] = 0u;
    }
    const int L = 32 * w + lane;
    const int x0 = L * XPL;                     // this lane's first token
    float left = (L == 0) ? 0.0f : kNeg;        // frame 0: v_prev(x=0) = 0, everything else -1e9
    const int ntiles = (ty + kTileY - 1) / kTileY;
    int stage = g0 % ring.nstages;
    uint32_t phase = (uint32_t)(g0 / ring.nstages) & 1u;
    for (int t = 0; t < ntiles; ++t) {
        const int gt = g0 + t;
        long long t0 = 0;
        if (wait_acc) t0 = clock64();
        MAS_TRACE(11 + w, gt);
        mbar_wait(&ring.full[stage], phase);
        if (w == 1) mbar_wait(&edge_full[gt & 3], (uint32_t)(gt >> 2) & 1u);   // warp 0 finished this tile
        if (wait_acc) *wait_acc += clock64() - t0;
        MAS_TRACE(7 + 2 * w, gt);
        const float *tile = ring.stages + stage * ring.stage_floats;
        float *etile = edge + ((gt & 3) << 5);
        const int y0 = t * kTileY;
        const int nsteps = min(kTileY, ty - y0);
        const bool diag = y0 < tx;  // some token x > y still exists in this tile
        if (nsteps == kTileY) {
            if (diag) dp_tile2<XPL, true, true, FMAX>(V, acc, left, tile, etile, w, lane, x0, y0, nsteps);
            else dp_tile2<XPL, false, true, FMAX>(V, acc, left, tile, etile, w, lane, x0, y0, nsteps);
        } else {
            dp_tile2<XPL, true, false, FMAX>(V, acc, left, tile, etile, w, lane, x0, y0, nsteps);
        }
        __syncwarp();
        MAS_TRACE(8 + 2 * w, gt);
        if (lane == 0) {
            mbar_arrive(&ring.empty[stage]);
            if (w == 0) mbar_arrive(&edge_full[gt & 3]);   // release: lane 31's stores are ordered by __syncwarp
        }
        if (++stage == ring.nstages) {
            stage = 0;
            phase ^= 1u;
        }
        // x == y always steps down (core.pyx:34 `index == y`), token 0 never does.
        if (diag) {
#pragma unroll
            for (int j = 0; j < XPL; ++j)
                if (((x0 + j) >> 5) == t) acc[j] |= 1u << ((x0 + j) & 31);
        }
        if (L == 0) acc[0] = 0u;
        uint32_t *dst = bits + (size_t)t * xrows + L;
        MAS_CHECK(((XPL - 1) << 6) + L < xrows && t < ntiles);
#pragma unroll
        for (int j = 0; j < XPL; ++j) {
            dst[j << 6] = acc[j];
            acc[j] = 0u;
        }
    }
    // total alignment score: token tx-1 = global lane (tx-1)/XPL, slot (tx-1)%XPL
    const int ql = (tx - 1) / XPL, qj = (tx - 1) - ql * XPL;
    *owns = (ql >> 5) == w;
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < XPL; ++j)
        if (j == qj) s = V[j];
    return __shfl_sync(kFull, s, ql & 31);
}

template <int XPLMAX, bool FMAX = false>
__device__ __forceinline__ float prior_forward2_dispatch(const TileRing &ring, uint32_t *bits, int xrows,
                                                         int tx, int ty, int lane, int w, int g0,
                                                         float *edge, uint64_t *edge_full, int *owns,
                                                         long long *wacc)
{
    const int xpl = (tx + 63) >> 6;
#define MAS_CASE2(N)                                                                                       \
    case N:                                                                                                \
        if constexpr (N <= XPLMAX)                                                                         \
            return dp_forward2<N, FMAX>(ring, bits, xrows, tx, ty, lane, w, g0, edge, edge_full, owns, wacc);    \
        break;
    switch (xpl) {
        MAS_CASE2(1) MAS_CASE2(2) MAS_CASE2(3) MAS_CASE2(4) MAS_CASE2(5) MAS_CASE2(6) MAS_CASE2(7) MAS_CASE2(8)
    default: break;
    }
#undef MAS_CASE2
    *owns = 0;
    return 0.0f;
}

// ------------------------------------------------------------------------------------
// Chain of DP warps across the CTAs of a thread-block cluster (long token axes: T_x > 256).
// CTA h of the cluster owns the tokens [xlo, xlo + xs) of the utterance (xs = 128 or 256) and runs them
// on two DP warps exactly like dp_forward2; the chain position of a warp is 2h + w.  Inside a CTA the
// boundary value travels through the `edge` ring as before.  Across CTAs it travels through
// DISTRIBUTED SHARED MEMORY: warp 1 of CTA h stores its last token's 32 values of a tile straight into
// a ring in CTA h+1's shared memory (st.shared::cluster) and arrives on that CTA's mbarrier; warp 0 of
// CTA h+1 runs behind it and hands the slot back with a remote arrive (kXRing tiles of slack).
// Tiles are counted globally (tile tt = frames 32 tt ..): CTA h only works on the tiles that hold band
// cells of its tokens, [t_lo, t_lo + ntiles), t_lo = xlo / 32.  Its first token needs V[xlo-1, 32 t_lo - 1],
// i.e. the LAST value of the left CTA's tile t_lo - 1: the exchange on a link covers the tiles
// t_lo - 1 .. in_last (in_last = last tile of the left CTA), the first one only seeds `left`.
// Direction words go to global memory (bits[tt * bstride + row], row local to the CTA).
// ------------------------------------------------------------------------------------
constexpr int kXRing = 8;   // tiles of boundary values in flight on a cluster link

struct ChainCtx {
    uint32_t in_ring;       // shared::cta address: [kXRing][32] floats written by the left CTA's warp 1
    uint64_t *in_full;      // [kXRing] count 1, arrived by the left CTA
    uint32_t left_empty;    // shared::cluster address of the left CTA's out_empty[0]
    uint32_t right_ring;    // shared::cluster address of the right CTA's in_ring
    uint32_t right_full;    // shared::cluster address of the right CTA's in_full[0]
    uint64_t *out_empty;    // [kXRing] count 1, arrived by the right CTA
    int n_in, n_out;        // boundary tiles received / sent so far (CTA lifetime)
};

template <int XPL, bool DIAG, bool FULL, bool FMAX, bool OUT_LOCAL>
__device__ __forceinline__ void dp_tile_chain(float (&V)[XPL], uint32_t (&acc)[XPL], float &left,
                                              const float *__restrict__ stage, uint32_t in_addr, uint32_t out_addr,
                                              int L, int lane, int x0, int y0, int nsteps)
{
    const uint32_t rowbase = smem_u32(stage + (L << 5));
    const int sw = lane & 7;
#pragma unroll 1
    for (int g = 0; g < 8; ++g) {
        const int s0 = g << 2;
        if (!FULL && s0 >= nsteps) break;
        float4 vv[XPL];
#pragma unroll
        for (int j = 0; j < XPL; ++j) vv[j] = lds128(rowbase + (j << 13) + ((g ^ sw) << 4));   // rows j*64 + L
        const float4 ev = in_addr ? lds128(in_addr + (g << 4)) : make_float4(kNeg, kNeg, kNeg, kNeg);
        float4 out;
        float v[XPL];
#pragma unroll
        for (int j = 0; j < XPL; ++j) v[j] = f4_get<0>(vv[j]);
        out.x = dp_step2<XPL, DIAG, FMAX>(V, acc, v, left, lane, x0, y0 + s0, 1u << s0, ev.x);
        out.y = out.z = out.w = 0.0f;
        if (FULL || s0 + 1 < nsteps) {
#pragma unroll
            for (int j = 0; j < XPL; ++j) v[j] = f4_get<1>(vv[j]);
            out.y = dp_step2<XPL, DIAG, FMAX>(V, acc, v, left, lane, x0, y0 + s0 + 1, 2u << s0, ev.y);
        }
        if (FULL || s0 + 2 < nsteps) {
#pragma unroll
            for (int j = 0; j < XPL; ++j) v[j] = f4_get<2>(vv[j]);
            out.z = dp_step2<XPL, DIAG, FMAX>(V, acc, v, left, lane, x0, y0 + s0 + 2, 4u << s0, ev.z);
        }
        if (FULL || s0 + 3 < nsteps) {
#pragma unroll
            for (int j = 0; j < XPL; ++j) v[j] = f4_get<3>(vv[j]);
            out.w = dp_step2<XPL, DIAG, FMAX>(V, acc, v, left, lane, x0, y0 + s0 + 3, 8u << s0, ev.w);
        }
        if (out_addr && lane == 31) {
            if (OUT_LOCAL)   // the boundary stays in this CTA (warp 0 -> warp 1)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(out_addr + (g << 4)), "f"(out.x), "f"(out.y),
                             "f"(out.z), "f"(out.w)
                             : "memory");
            else
                st_cluster_f4(out_addr + (g << 4), out.x, out.y, out.z, out.w);
        }
    }
}

// Forward pass of the tokens [xlo, xlo + 64 XPL) of one utterance by DP warp `w` (0/1) of this CTA.
// `in_last`  (warp 0): last tile for which the left CTA publishes a boundary (< t_lo - 1: no left neighbour);
// `out_first` (warp 1): first tile whose boundary the right CTA wants (> last tile: no right neighbour).
// Returns V[t_x-1, t_y-1] when this warp owns token t_x-1 (`*owns`).
template <int XPL, bool FMAX = false>
__device__ __noinline__ float dp_forward_chain(const TileRing ring, uint32_t *bits, int bstride, int tx, int ty,
                                               int lane, int w, int g0, float *edge, uint64_t *edge_full,
                                               int xlo, int t_lo, int ntiles, int in_last, int out_first,
                                               ChainCtx &cx, int *owns, long long *wacc = nullptr)
{
    float V[XPL];
    uint32_t acc[XPL];
#pragma unroll
    for (int j = 0; j < XPL; ++j) {
        V[j] = kNeg;
        acc[j] = 0u;
    }
    const int L = 32 * w + lane;
    const int x0 = xlo + L * XPL;               // this lane's first token (global index)
    float left = (x0 == 0) ? 0.0f : kNeg;       // frame 0: v_prev(x=0) = 0, everything else -1e9
    if (w == 0 && in_last >= t_lo - 1) {
        // V[xlo-1, 32 t_lo - 1]: the last value of the left CTA's tile t_lo - 1
        const int slot = cx.n_in % kXRing;
        mbar_wait_cluster(&cx.in_full[slot], (uint32_t)(cx.n_in / kXRing) & 1u);
        float e;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(e) : "r"(cx.in_ring + (uint32_t)slot * 128u + 124u));
        if (lane == 0) left = e;
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(cx.left_empty + (uint32_t)slot * 8u);
        ++cx.n_in;
    }
    int stage = g0 % ring.nstages;
    uint32_t phase = (uint32_t)(g0 / ring.nstages) & 1u;
    for (int t = 0; t < ntiles; ++t) {
        const int gt = g0 + t, tt = t_lo + t;
        long long c0 = 0, c1 = 0;
        if (wacc) c0 = clock64();
        mbar_wait(&ring.full[stage], phase);
        if (wacc) c1 = clock64();
        uint32_t in_addr = 0u, out_addr = 0u;
        int in_slot = -1, out_slot = -1;
        if (w == 1) {
            mbar_wait(&edge_full[gt & 3], (uint32_t)(gt >> 2) & 1u);   // warp 0 finished this tile
            in_addr = smem_u32(edge + ((gt & 3) << 5));
            if (tt >= out_first) {
                out_slot = cx.n_out % kXRing;
                if (cx.n_out >= kXRing)
                    mbar_wait_cluster(&cx.out_empty[out_slot], (uint32_t)(cx.n_out / kXRing - 1) & 1u);
                out_addr = cx.right_ring + (uint32_t)out_slot * 128u;
                ++cx.n_out;
            }
        } else {
            out_addr = smem_u32(edge + ((gt & 3) << 5));
            if (tt <= in_last) {
                in_slot = cx.n_in % kXRing;
                mbar_wait_cluster(&cx.in_full[in_slot], (uint32_t)(cx.n_in / kXRing) & 1u);
                in_addr = cx.in_ring + (uint32_t)in_slot * 128u;
                ++cx.n_in;
            }
        }
        if (wacc) {   // profiling aid: starved of tiles / waiting for a boundary (neighbour warp or CTA)
            wacc[0] += c1 - c0;
            wacc[1] += clock64() - c1;
        }
        const float *tile = ring.stages + stage * ring.stage_floats;
        const int y0 = tt * kTileY;
        const int nsteps = min(kTileY, ty - y0);
        const bool diag = y0 < tx;  // some token x > y may still exist in this tile
        if (w == 0) {
            if (nsteps == kTileY) {
                if (diag) dp_tile_chain<XPL, true, true, FMAX, true>(V, acc, left, tile, in_addr, out_addr, L, lane, x0, y0, nsteps);
                else dp_tile_chain<XPL, false, true, FMAX, true>(V, acc, left, tile, in_addr, out_addr, L, lane, x0, y0, nsteps);
            } else {
                dp_tile_chain<XPL, true, false, FMAX, true>(V, acc, left, tile, in_addr, out_addr, L, lane, x0, y0, nsteps);
            }
        } else {
            if (nsteps == kTileY) {
                if (diag) dp_tile_chain<XPL, true, true, FMAX, false>(V, acc, left, tile, in_addr, out_addr, L, lane, x0, y0, nsteps);
                else dp_tile_chain<XPL, false, true, FMAX, false>(V, acc, left, tile, in_addr, out_addr, L, lane, x0, y0, nsteps);
            } else {
                dp_tile_chain<XPL, true, false, FMAX, false>(V, acc, left, tile, in_addr, out_addr, L, lane, x0, y0, nsteps);
            }
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&ring.empty[stage]);
            if (w == 0) mbar_arrive(&edge_full[gt & 3]);
            if (in_slot >= 0) mbar_arrive_cluster(cx.left_empty + (uint32_t)in_slot * 8u);   // slot read: hand it back
        }
        // the thread that stored the boundary values publishes them (release at cluster scope)
        if (out_slot >= 0 && lane == 31) mbar_arrive_cluster(cx.right_full + (uint32_t)out_slot * 8u);
        if (++stage == ring.nstages) {
            stage = 0;
            phase ^= 1u;
        }
        // x == y always steps down (core.pyx:34 `index == y`), token 0 never does.
        if (diag) {
#pragma unroll
            for (int j = 0; j < XPL; ++j)
                if (((x0 + j) >> 5) == tt) acc[j] |= 1u << ((x0 + j) & 31);
        }
        if (x0 == 0) acc[0] = 0u;
        uint32_t *dst = bits + (size_t)tt * bstride + L;
        MAS_CHECK(tt >= 0 && tt * kTileY < ty && xlo + ((XPL - 1) << 6) + L < bstride);
#pragma unroll
        for (int j = 0; j < XPL; ++j) {
            dst[j << 6] = acc[j];
            acc[j] = 0u;
        }
    }
    MAS_CHECK(cx.n_in >= 0 && cx.n_out >= 0);
    // total alignment score: token tx-1 = local lane (tx-1-xlo)/XPL, slot (tx-1-xlo)%XPL
    const int xl = tx - 1 - xlo;
    const int ql = xl / XPL, qj = xl - ql * XPL;
    *owns = (xl >= 0 && xl < 64 * XPL && (ql >> 5) == w) ? 1 : 0;
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < XPL; ++j)
        if (j == qj) s = V[j];
    return __shfl_sync(kFull, s, ql & 31);
}

template <int XPLMAX, bool FMAX = false>
__device__ __forceinline__ float chain_forward_dispatch(const TileRing &ring, uint32_t *bits, int bstride, int tx,
                                                        int ty, int txl, int lane, int w, int g0, float *edge,
                                                        uint64_t *edge_full, int xlo, int t_lo, int ntiles,
                                                        int in_last, int out_first, ChainCtx &cx, int *owns,
                                                        long long *wacc = nullptr)
{
    const int xpl = (txl + 63) >> 6;
#define MAS_CASEC(N)                                                                                              \
    case N:                                                                                                       \
        if constexpr (N <= XPLMAX)                                                                                \
            return dp_forward_chain<N, FMAX>(ring, bits, bstride, tx, ty, lane, w, g0, edge, edge_full, xlo, t_lo, \
                                             ntiles, in_last, out_first, cx, owns, wacc);                         \
        break;
    switch (xpl) {
        MAS_CASEC(1) MAS_CASEC(2) MAS_CASEC(3) MAS_CASEC(4)
    default: break;
    }
#undef MAS_CASEC
    *owns = 0;
    return 0.0f;
}

// Backtrack over direction words kept in GLOBAL memory, by a whole warp.  One dependent load per word from
// L2 (~700 cycles) would dominate a long utterance, so the warp keeps a WINDOW of kBtWinC chunks x 32 tokens
// in shared memory (`win`, 32 * kBtWinC words): the 32 lanes fetch it with cp.async, all requests in flight at
// once (one L2 latency per window), and the walk then costs one broadcast LDS per step; every lane follows
// the same walk.  Tokens are the CTA's local ones (global token idx = xlo + local), rows by RowMap(txl, 6).
// Starts on token `idx` at frame `y` (the last frame of that token); returns the last frame of token
// xlo - 1 when the walk leaves the CTA's tokens downwards (the left CTA continues there), or -1 when it
// ended on token 0.
constexpr int kBtWinC = 8;
__device__ __forceinline__ int backtrack_bits_window(const uint32_t *bits, int bstride, int txl, int xlo, int idx,
                                                     int y, int *first, int *dur, uint32_t *win, int lane)
{
    const RowMap rm(txl, 6);
    const uint32_t swin = smem_u32(win);
    int top = y, c = y >> 5;
    int wx = -1, wc = -1;
    while (y >= 0 && idx != 0) {
        const int xl = idx - xlo;
        if (wx < 0 || xl < wx - 31 || c < wc - (kBtWinC - 1)) {
            __syncwarp();          // every lane is done reading the old window
            wx = xl;
            wc = c;
            const int mx = wx - lane;
            const int row = mx >= 0 ? rm.row(mx) : 0;
#pragma unroll
            for (int n = 0; n < kBtWinC; ++n) {
                const bool ok = mx >= 0 && wc - n >= 0;
                cp_async4(win + lane * kBtWinC + n, ok ? bits + (size_t)(wc - n) * bstride + row : bits, ok ? 4u : 0u);
            }
            asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
            __syncwarp();
        }
        uint32_t w;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(swin + 4u * (uint32_t)((wx - xl) * kBtWinC + (wc - c))));
        const int s = y & 31;
        const uint32_t m = w & (0xffffffffu >> (31 - s));
        if (m == 0u) {  // stays on this token down to the chunk start
            y = (c << 5) - 1;
            --c;
            continue;
        }
        const int p = 31 - __clz(m);
        const int ys = (c << 5) + p;
        MAS_CHECK(xl >= 0 && xl < txl && ys <= top);
        if (lane == 0) {
            first[idx] = ys;
            dur[idx] = top - ys + 1;
        }
        --idx;
        y = ys - 1;
        top = y;
        if (idx < xlo) return y;
        if (p == 0) --c;   // crossed into the previous chunk
    }
    if (top >= 0 && lane == 0) {
        first[idx] = 0;
        dur[idx] = top + 1;
    }
    return -1;
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

template <int XPLMAX, bool FMAX = false>
__device__ __forceinline__ float prior_forward_dispatch(const TileRing &ring, uint32_t *bits,
                                                        int xrows, int tx, int ty, int lane, int g0,
                                                        long long *wacc)
{
    const int xpl = (tx + 31) >> 5;
#define MAS_CASE(N)                                                                           \
    case N:                                                                                   \
        if constexpr (N <= XPLMAX) return dp_forward<N, FMAX>(ring, bits, xrows, tx, ty, lane, g0, wacc); \
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


}  // namespace mas
