// mas_dp3.cuh -- the MAS recurrence with SKEWED LANES: one warp, no shuffle on the critical path.
//
// Reference semantics (paths relative to /root/reference/): forward core.pyx:17-30, backtrack
// core.pyx:15,32-35.  Same results as mas_dp.cuh, different schedule and a cheaper cell:
//
//   * lane r owns the XPL consecutive tokens x = r*XPL + j and, at warp iteration tau, works on frame
//     f = tau - r: every lane runs ONE FRAME BEHIND its left neighbour.  The only cross-lane value of
//     the recurrence, V[x-1, f-1] for the lane's first token, was therefore produced by lane r-1 two
//     iterations earlier; its shuffle is issued a whole iteration before it is consumed.  In the
//     lock-step formulation (mas_dp.cuh) the same shuffle sits on a diagonal dependency chain and an
//     in-order warp stalls on it with nothing else to issue (ncu, profiles/r1: 73 cycles per frame for
//     3 tokens per lane, 70 % of them stall cycles).
//   * the values of iteration tau+2 are loaded (LDS.32, conflict-free by layout) while iteration tau
//     is computed, so shared-memory latency is off the chain as well.
//   * what is left is ISSUE-bound on the 16-lane ALU pipe (FSETP / FSEL / FMNMX / LOP3 / SHF occupy it
//     two cycles per warp instruction; FADD runs on the FMA pipe).  The cell is therefore
//         d = cur - up            FADD   (up > cur  <=>  cur - up < 0: exact without NaN; V is never -0)
//         acc = (acc << 1) | d>>31  SHF  (funnel shift takes the sign bit straight out of d)
//         V = fmax(up, cur) + v   FMNMX + FADD
//     = 2 ALU + 2 FMA-pipe instructions, against FSETP + FSEL + SEL/LOP3 + FADD (+ guards) before.
//     fmax(up, cur) is the reference's (up > cur) ? up : cur on every input without NaN.
//   * no per-cell guards: the PRODUCER stores 0.0 for every cell above the diagonal (x > f), so those
//     cells stay exactly -1e9 (max(-1e9, -1e9) + 0) as the reference's `x == y ? -1e9` needs; only
//     block 0 (whose lanes also see frames < 0) selects, and the last blocks capture V at f == t_y-1
//     (the score) instead of freezing the recurrence.
//   * direction words are accumulated MSB first; at the end of a 32-iteration block the frame-aligned
//     word of chunk g-1 is funnelshift_r(brev(previous), brev(current), lane).
//
// The schedule is stated in plain Python in tests/dp3_model.py and checked there against the oracle
// (tests/test_dp3_model.py); this file is its transcription.
//
// Ring layout: row x (natural token order) is a circular buffer of RC frames (128 or 96), row pitch
// RC + 4 floats; frame f of the utterance whose first tile is tile g0 of the ring's lifetime sits at
// float column (32*g0 + f) mod RC -- the same column in every row, so producers write aligned 16-byte
// chunks (cp.async / STS.128; thread-per-row STS.128 of consecutive rows is conflict-free thanks to the
// +4 pad) and utterances of different lengths can follow each other in one ring.  The skew lives in the
// CONSUMER's address: lane r reads column (32*g0 + tau - r) mod RC of its own rows, a per-lane register
// advanced by 4 bytes per iteration; bank = 4 (r XPL + j) + tau - r = (4 XPL - 1) r + const (mod 32),
// and 4 XPL - 1 is odd: the 32 lanes always hit 32 distinct banks (profiles/ring3_banks.py).
#pragma once

#include "mas_dp.cuh"

namespace mas {

// frames per row ring: 128 (4 tiles, a mask wraps the column) or 96 (3 tiles, where shared memory is
// short -- the fused tensor-core kernel: compare + select wraps it); row pitch = ring + 4 floats
constexpr int kRing3Cols = 128;
constexpr int kRing3Pitch = kRing3Cols + 4;
constexpr int kRing3Stages = kRing3Cols / kTileY;

template <int RC>
struct Ring3Geom {
    static_assert(RC == 128 || RC == 96, "ring of 3 or 4 tiles");
    static constexpr int cols = RC, pitch = RC + 4, stages = RC / kTileY;
    __device__ __forceinline__ static int stage(int g) { return RC == 128 ? (g & 3) : (g % 3); }
    __device__ __forceinline__ static uint32_t parity(int g) { return (uint32_t)(RC == 128 ? (g >> 2) : (g / 3)) & 1u; }
    __device__ __forceinline__ static int wrap(int c) { return RC == 128 ? (c & 127) : (c % 96); }   // c >= 0
};

// x / xpl for x < 512, xpl <= 16 (lane that owns token x)
struct XDiv {
    uint32_t inv;
    __device__ __forceinline__ explicit XDiv(int xpl) : inv((65536u + xpl - 1) / xpl) {}
    __device__ __forceinline__ int operator()(int x) const { return (int)((uint32_t)x * inv >> 16); }
};

// float column (within every row) of lifetime frame fl (= 32*g0 + f >= 0)
template <int RC = kRing3Cols>
__device__ __forceinline__ int dp3_col(int fl) { return Ring3Geom<RC>::wrap(fl); }

struct Ring3 {
    float *rows;        // xrows x pitch floats, 16-byte aligned
    uint64_t *full;     // [stages] producers -> DP warp, one per tile
    uint64_t *empty;    // [stages] DP warp -> producers
};

__device__ __forceinline__ float lds32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// values of one iteration for this lane's XPL tokens (rows x0 .. x0+XPL-1, same column)
// `t4` = byte column, already wrapped into the ring.  The destination is declared read-write: the load
// then cannot be hoisted above the last use of the value it replaces, so ptxas reuses the register in
// place instead of loading into a fresh one and MOV-ing it over right away -- a MOV that waits out the
// whole shared-memory latency once per loop trip (ncu: the top short_sb stall of the first version).
template <int XPL, int RC>
__device__ __forceinline__ void dp3_load(float (&v)[XPL], uint32_t rowbase, uint32_t t4)
{
    MAS_CHECK(t4 < (uint32_t)(RC * 4) && (t4 & 3u) == 0u);
    const uint32_t a = rowbase + t4;
#pragma unroll
    for (int j = 0; j < XPL; ++j)
        asm volatile("ld.shared.f32 %0, [%1];" : "+f"(v[j]) : "r"(a + (uint32_t)(j * (RC + 4) * 4)));
}
template <int RC>
__device__ __forceinline__ uint32_t dp3_next(uint32_t t4)
{
    if (RC == 128) return (t4 + 4u) & 508u;
    return (t4 == (uint32_t)(RC * 4 - 4)) ? 0u : t4 + 4u;
}

// 32 iterations (one block).  `f` = this lane's frame at the block's first iteration (tau - lane).
// EDGE blocks: `head` (block 0) selects "x > f -> -1e9" (also covers f < 0, where the ring holds stale
// data); every EDGE block captures V at f == ty - 1 into sv.  `wait_bar` (may be null): barrier of the
// NEXT tile, waited before the last 8 iterations -- their prefetches are the first loads that touch it.
// kDp3Unroll iterations, fully unrolled.  Kept SHORT on purpose: the L0 instruction cache of a
// sub-partition is ~6 KB and the fused kernels run an epilogue warp's tile loop next to this one
// (ncu on the first version, 8 iterations = 4.6 KB at 6 tokens per lane: 62 % of the DP warp's samples
// were stall_no_inst, profiles/r2_tc_dp3_icache.txt)
constexpr int kDp3Unroll = 2;
template <int XPL, bool EDGE, int RC>
__device__ __forceinline__ void dp3_itern(float (&V)[XPL], uint32_t (&acc)[XPL], float (&buf)[2][XPL],
                                          float (&sv)[XPL], float &left, uint32_t &t4, uint32_t rowbase,
                                          int lane, int x0, int &f, int ty, bool head)
{
#pragma unroll
    for (int u = 0; u < kDp3Unroll; ++u) {
        // state at the end of the previous iteration: consumed one iteration from now
        const float cap = __shfl_up_sync(kFull, V[XPL - 1], 1);
        const float(&v)[XPL] = buf[u & 1];
#pragma unroll
        for (int j = XPL - 1; j >= 0; --j) {
            const float up = (j == 0) ? left : V[j > 0 ? j - 1 : 0];   // V[x-1, f-1]
            const float d = __fsub_rn(V[j], up);                       // < 0  <=>  up > cur (core.pyx:30)
            acc[j] = __funnelshift_l(__float_as_uint(d), acc[j], 1);
            float nv = __fadd_rn(fmaxf(up, V[j]), v[j]);
            if (EDGE) {
                nv = (head && x0 + j > f) ? kNeg : nv;                 // not reachable yet (or f < 0)
                sv[j] = (f == ty - 1) ? nv : sv[j];
            }
            V[j] = nv;
        }
        left = (lane == 0) ? kNeg : cap;   // token 0: v_prev = -1e9 after frame 0 (core.pyx:23-27)
        dp3_load<XPL, RC>(buf[u & 1], rowbase, t4);   // iteration tau + 2
        t4 = dp3_next<RC>(t4);
        if (EDGE) ++f;
    }
}

template <int XPL, bool EDGE, int RC>
__device__ __forceinline__ void dp3_block(float (&V)[XPL], uint32_t (&acc)[XPL], float (&buf)[2][XPL],
                                          float (&sv)[XPL], float &left, uint32_t &t4, uint32_t rowbase,
                                          int lane, int x0, int f, int ty, bool head, uint64_t *wait_bar,
                                          uint32_t wait_parity, long long *wait_acc)
{
#pragma unroll 1
    for (int k = 0; k < (kTileY - 8) / kDp3Unroll; ++k)
        dp3_itern<XPL, EDGE, RC>(V, acc, buf, sv, left, t4, rowbase, lane, x0, f, ty, head);
    if (wait_bar) {
        if (wait_acc) {   // profiling aid: cycles this warp spends starved of tiles
            const long long t0 = clock64();
            mbar_wait(wait_bar, wait_parity);
            *wait_acc += clock64() - t0;
        } else {
            mbar_wait(wait_bar, wait_parity);
        }
    }
#pragma unroll 1
    for (int k = 0; k < 8 / kDp3Unroll; ++k)
        dp3_itern<XPL, EDGE, RC>(V, acc, buf, sv, left, t4, rowbase, lane, x0, f, ty, head);
}

// frame-aligned direction words of chunk c from the previous (already bit-reversed) and the current
// (MSB-first) block accumulators
template <int XPL>
__device__ __forceinline__ void dp3_flush(uint32_t *bits, int xrows, int nch, int c, int lane, int x0,
                                          uint32_t (&prev)[XPL], uint32_t (&acc)[XPL])
{
#pragma unroll
    for (int j = 0; j < XPL; ++j) {
        const uint32_t cur = __brev(acc[j]);
        uint32_t w = __funnelshift_r(prev[j], cur, lane);
        const int x = x0 + j;
        if ((x >> 5) == c) w |= 1u << (x & 31);   // x == y always steps down (core.pyx:34 `index == y`)
        if (x == 0) w = 0u;                        // token 0 never does
        MAS_CHECK(x < xrows);
        if (c >= 0 && c < nch) bits[(size_t)c * xrows + x] = w;
        prev[j] = cur;
        acc[j] = 0u;
    }
}

// Forward pass of one utterance by ONE warp (1 <= t_x <= t_y, XPL == ceil(t_x / 32) <= 8).  Consumes
// tiles g0 .. g0+ceil(t_y/32)-1 of the ring's lifetime -- whose producers store 0.0 above the diagonal --
// writes bits[chunk * xrows + x] (natural token order) and returns V[t_x-1, t_y-1] in every lane.
template <int XPL, int RC = kRing3Cols>
__device__ __noinline__ float dp3_forward(const Ring3 ring, uint32_t *bits, int xrows, int tx, int ty,
                                          int lane, int g0, long long *wait_acc = nullptr)
{
    using G = Ring3Geom<RC>;
    float V[XPL], buf[2][XPL], sv[XPL];
    uint32_t acc[XPL], prev[XPL];
#pragma unroll
    for (int j = 0; j < XPL; ++j) {
        V[j] = kNeg;
        sv[j] = 0.0f;
        buf[0][j] = buf[1][j] = 0.0f;
        acc[j] = 0u;
        prev[j] = 0u;
    }
    const int x0 = lane * XPL;
    const int ntiles = (ty + kTileY - 1) / kTileY;   // == nch
    const int nblk = (ty + 31 + kTileY - 1) / kTileY;
    const uint32_t rowbase = smem_u32(ring.rows) + (uint32_t)(x0 * G::pitch * 4);
    uint32_t t4 = (uint32_t)(G::wrap(32 * g0 + RC - lane) * 4);   // byte column of iteration 0: frame -lane
    float left = (lane == 0) ? 0.0f : kNeg;     // frame 0: v_prev(x=0) = 0, everything else -1e9
    {
        long long t0 = 0;
        if (wait_acc) t0 = clock64();
        mbar_wait(&ring.full[G::stage(g0)], G::parity(g0));
        if (wait_acc) *wait_acc += clock64() - t0;
    }
    dp3_load<XPL, RC>(buf[0], rowbase, t4);
    t4 = dp3_next<RC>(t4);
    dp3_load<XPL, RC>(buf[1], rowbase, t4);
    t4 = dp3_next<RC>(t4);
    for (int g = 0; g < nblk; ++g) {
        const bool head = g == 0;
        const bool tail = 32 * g + 31 >= ty - 1;
        const int gt = g0 + g + 1;
        uint64_t *wb = (g + 1 < ntiles) ? &ring.full[G::stage(gt)] : nullptr;
        const uint32_t wp = G::parity(gt);
        if (head || tail)
            dp3_block<XPL, true, RC>(V, acc, buf, sv, left, t4, rowbase, lane, x0, 32 * g - lane, ty, head, wb, wp,
                                     wait_acc);
        else
            dp3_block<XPL, false, RC>(V, acc, buf, sv, left, t4, rowbase, lane, x0, 32 * g - lane, ty, false, wb,
                                      wp, wait_acc);
        dp3_flush<XPL>(bits, xrows, ntiles, g - 1, lane, x0, prev, acc);
        if (g >= 1) {   // every lane is past frame 32g: tile g-1 is free
            __syncwarp();
            if (lane == 0) mbar_arrive(&ring.empty[G::stage(g0 + g - 1)]);
        }
    }
    dp3_flush<XPL>(bits, xrows, ntiles, nblk - 1, lane, x0, prev, acc);   // t_y = 1 (mod 32)
    __syncwarp();
    if (lane == 0)
        for (int t = max(nblk - 1, 0); t < ntiles; ++t) mbar_arrive(&ring.empty[G::stage(g0 + t)]);
    const int ql = (tx - 1) / XPL, qj = (tx - 1) - ql * XPL;
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < XPL; ++j)
        if (j == qj) s = sv[j];
    return __shfl_sync(kFull, s, ql);
}

template <int XPLMAX, int RC = kRing3Cols>
__device__ __forceinline__ float dp3_forward_dispatch(const Ring3 &ring, uint32_t *bits, int xrows, int tx,
                                                      int ty, int lane, int g0, long long *wacc = nullptr)
{
    const int xpl = (tx + 31) >> 5;
#define MAS_CASE3(N)                                                                                 \
    case N:                                                                                          \
        if constexpr (N <= XPLMAX) return dp3_forward<N, RC>(ring, bits, xrows, tx, ty, lane, g0, wacc); \
        break;
    switch (xpl) {
        MAS_CASE3(1) MAS_CASE3(2) MAS_CASE3(3) MAS_CASE3(4) MAS_CASE3(5) MAS_CASE3(6) MAS_CASE3(7) MAS_CASE3(8)
    default: break;
    }
#undef MAS_CASE3
    return 0.0f;
}

// Backtrack over direction words in NATURAL token order (bits[chunk * xrows + x], shared memory), one
// lane (core.pyx:32-35).  Jumps from token boundary to token boundary: inside a 32-frame word the next
// decrement is the highest set bit at or below the current frame.  Only the FIRST FRAME of every token
// is recorded (first[x]; durations follow in parallel: dur[x] = first[x+1] - first[x]), addresses are
// running pointers, and the word of the next token is prefetched, so one transition is the dependent
// chain mask -> AND -> FLO -> add (~25 cycles; the round-1 walk cost 165 per token).
// Requires 1 <= t_x <= t_y (every token gets >= 1 frame).
__device__ __forceinline__ void backtrack_nat(const uint32_t *bits, int xrows, int tx, int ty, int *first)
{
    int idx = tx - 1;
    int c = (ty - 1) >> 5, s = (ty - 1) & 31;
    const uint32_t rowb = (uint32_t)xrows * 4u;
    uint32_t pw = smem_u32(bits) + 4u * (uint32_t)(c * xrows + idx);   // &bits[c][idx]
    auto ld = [](uint32_t a) {
        uint32_t w;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(a));
        return w;
    };
    uint32_t w = ld(pw);
    uint32_t wn = idx > 0 ? ld(pw - 4u) : 0u;
    while (idx > 0) {
        const uint32_t m = w & (0xffffffffu >> (31 - s));
        if (m == 0u) {            // stays on this token down to the chunk start
            if (c == 0) break;    // (cannot happen for idx > 0 when t_x <= t_y: bit idx of chunk idx>>5 is set)
            --c;
            pw -= rowb;
            s = 31;
            w = ld(pw);
            wn = ld(pw - 4u);
            continue;
        }
        const int p = 31 - __clz(m);
        MAS_CHECK(idx > 0 && idx < tx && c >= 0 && (c << 5) + p < ty);
        first[idx] = (c << 5) + p;
        --idx;
        pw -= 4u;
        w = wn;
        if (p == 0) {             // crossed into the previous chunk
            --c;                  // c >= 1 here: frame 0 belongs to token 0 and idx was > 0
            pw -= rowb;
            s = 31;
            w = ld(pw);
        } else {
            s = p - 1;
        }
        wn = idx > 0 ? ld(pw - 4u) : 0u;
    }
    first[0] = 0;
}

}  // namespace mas
