// mas_prior_tc2.cu -- fused Gaussian log-prior + MAS, tensor-core prior, DP STRAIGHT FROM TENSOR MEMORY.
//
// Same contract and the same front half as mas_prior_tc.cu (3xTF32 tcgen05.mma, mu_x resident in
// TMEM as the A operand, K-major y slabs as B), but the accumulator never goes through shared
// memory: the four warps that may read TMEM (one per 32-lane quarter) ARE the DP warps.  Token x
// lives in TMEM lane L = x / P of M tile j = x % P (P = 1 or 2 tokens per lane), so lane L of
// quarter-warp q owns consecutive tokens and reads its own accumulator rows with tcgen05.ld: a
// row is 32 frames of one token, exactly what the frame-sequential recurrence consumes.  The
// x-1 neighbour is a register or one shuffle; the value that crosses a warp boundary goes
// through a small ring in shared memory, each warp running 8 frames behind its left neighbour
// (one mbarrier hand-off per 8 frames).  No epilogue warps, no tile ring, no transposition.
#include <algorithm>
#include <type_traits>

#include "mas_dp.cuh"
#include "mas_internal.h"

namespace mas {

namespace {

constexpr int kTcThreads = 384;
// warps 0-3 epilogue, 8-11 mu_x movers (TMEM lane quarter = warp % 4 is a hardware rule); the rest is
// placed by scheduler (warp % 4): the DP warp shares its scheduler with the latency-bound backtrack
// warp, not with a loader
constexpr int kTcMma = 5, kTcLoader = 6, kTcLoader2 = 7, kTcBack = 4;   // warps 0-3: DP (TMEM quarters), 8-11: movers
constexpr int kTcSlabs = 4;       // y slabs (hi+lo) in flight
constexpr int kTcYsq = 8;         // ring of per-slab -0.5|y|^2 vectors (> kTcSlabs + D buffers)
constexpr int kTcStage = 2;        // staging buffers [F][32 frames] behind the cp.async copies: one per loader warp
constexpr int kTcMsq = 4;         // ring of per-utterance -0.5|mu|^2 vectors (> utterances an accumulator lags)
constexpr int kTcZeroBytes = 4096;
constexpr int kTcTmemCols = 512;
#ifndef MAS_EXP_PASS0
#define MAS_EXP_PASS0 0
#endif
constexpr int kExpPass0 = MAS_EXP_PASS0;
#ifdef MAS_TC2_DPSTATS   // instrumented build (profiles/prior_tc_stats.py --dp): where the DP warps wait
#define DPSTAT(x) x
#else
#define DPSTAT(x)
#endif   // experiment: skip the first passes (wrong results, timing only)
constexpr int kTc2BarBytes = 3072;   // barriers (1 KB) + DP edge rings (3 x 16 x 8 floats) + their barriers

// ---------------------------------------------------------------- tcgen05 wrappers
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]; `acc` = 0 overwrites D
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                          uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
                 "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
                 : "memory");
}
// split descriptor + provably uniform operands: see mas_prior_tc.cu (one instruction per MMA for the issuer)
__device__ __forceinline__ void tc_mma_ts2(uint32_t d_tmem, uint32_t a_tmem, uint32_t bdesc_lo, uint32_t bdesc_hi,
                                           uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 bd, {%2, %3};\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t}" ::"r"(d_tmem),
                 "r"(a_tmem), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ uint32_t uniform_u32(uint32_t v) { return __reduce_min_sync(kFull, v); }

// B operand = one 32-frame slab of y, K-MAJOR in shared memory, SWIZZLE_32B: per k step (8
// features) a [32 frames][8 features] block of 32-byte rows, 1 KB contiguous; the two 16-byte
// halves of row j are swapped when (j >> 2) & 1.  SBO = 256 B between groups of 8 rows.
// Measured on B200 (profiles/microbench/tc_prior.cu, 128x32x8 tf32, A in TMEM, 4 accumulators in
// flight): 33 cycles/MMA with this layout, 45-190 with SWIZZLE_128B rows, 135 with the MN-major
// layout y has in HBM -- hence the transposing loader.
__device__ __forceinline__ uint64_t tc_bdesc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)(16 >> 4) << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version of sm_100
    d |= (uint64_t)6 << 61;  // SWIZZLE_32B
    return d;
}
// one elected lane of a converged warp (operands stay warp-uniform for the tcgen05 issue)
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ float tf32_rn(float x)
{
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
// coarse cycle accounting (PriorTcArgs::stats, MAS_PRIOR_STATS=1; profiles/prior_tc_stats.py)
struct TcStat {
    long long acc = 0, t0 = 0;
    bool on;
    __device__ __forceinline__ explicit TcStat(bool e) : on(e) {}
    __device__ __forceinline__ void begin() { if (on) t0 = clock64(); }
    __device__ __forceinline__ void end() { if (on) acc += clock64() - t0; }
};

}  // namespace

// Shared-memory carve-up and TMEM column map of the tensor-core kernel, or ok == 0 when the
// shape does not qualify (T_x > 256, F > 96, or the buffers do not fit): the CUDA-core kernel
// of mas_prior.cu takes those.
TcLayout tc2_layout(int F, int T_x, int T_y)
{
    TcLayout L{};
    L.ok = 0;
    L.Fp = (F + 7) / 8 * 8;
    L.xrows = 256;                             // direction-bit rows: token x -> row (x % P) * 128 + x / P
    L.nch = (T_y + 31) / 32;
    if (T_x > 256 || L.Fp > 96) return L;
    const int mt = (T_x + 127) / 128;          // M tiles of 128 tokens
    L.col_ahi = 0;
    L.col_alo = mt * L.Fp;
    L.col_d = 2 * mt * L.Fp;
    L.nb = std::min(4, (kTcTmemCols - L.col_d) / 64);   // accumulator buffers of 2 x 32 columns
    if (L.nb < 2) return L;
    const size_t bits = (size_t)L.nch * L.xrows * 4;
    const size_t slabs = (size_t)kTcSlabs * 2 * L.Fp * 128;   // hi + lo, K-major: 1 KB per k step of 8 features
    const size_t misc = (size_t)kTcYsq * 128 + (size_t)kTcMsq * 256 * 4 + (((size_t)T_x * 8 + 15) & ~(size_t)15) + 16 +
                        kTcZeroBytes + kTc2BarBytes + (size_t)kTcStage * L.Fp * 128;
    L.nstages = 0;                             // no tile ring: the DP reads the accumulators in TMEM
    for (int slots = 2; slots >= 1 && !L.ok; --slots)
        if (slots * bits + slabs + misc <= (size_t)kSmemBudget) {
            L.bits_slots = slots;
            L.ok = 1;
        }
    if (!L.ok) return L;
    L.off_stages = 0;
    L.off_slabs = 0;
    L.off_staging = L.off_slabs + slabs;
    L.off_bits = L.off_staging + (size_t)kTcStage * L.Fp * 128;
    L.off_ysq = L.off_bits + (size_t)L.bits_slots * bits;
    L.off_musq = L.off_ysq + (size_t)kTcYsq * 128;
    L.off_first = L.off_musq + (size_t)kTcMsq * 256 * 4;
    L.off_dur = L.off_first + (size_t)T_x * 4;
    L.off_zero = (L.off_dur + (size_t)T_x * 4 + 15) & ~(size_t)15;
    L.off_bars = (L.off_zero + kTcZeroBytes + 127) & ~(size_t)127;
    L.total = L.off_bars + kTc2BarBytes;
    // one CTA per SM: each CTA allocates all 512 TMEM columns
    L.total = std::max(L.total, (size_t)(kSmemBudget / 2 + 1024));
    return L;
}

// ------------------------------------------------------------------------------------
// The recurrence of mas_dp.cuh on accumulator rows read from TMEM.  One tile = 32 frames in four
// groups of 8: tcgen05.ld 8 columns per owned token, wait for the left neighbour warp's boundary
// values of the same 8 frames, run dp_step2 on them, publish this warp's own boundary values.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}

struct EdgeLink {            // hand-off between neighbouring DP warps, 16 groups of 8 frames deep
    uint32_t in, out;        // shared addresses of [16][8] boundary values: read / written by this warp
    uint32_t bar_in, bar_out;   // [16] mbarriers, count 1
};

__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra W;\n\t}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// One group of 8 frames on the accumulator columns `c` (already in registers).
// The DP warps are the latency-bound agents of this kernel (one dependent chain per warp), so the
// step is written for a short chain and few instructions.  Per frame and token: two FADDs for the
// prior (independent filler), FMNMX + FADD for the value (max(up, cur) equals the reference's
// `up > cur ? up : cur` on finite values), FSETP + predicated LOP3 for the direction bit; one
// SHFL + FSEL per frame for the left neighbour.
template <int P, bool DIAG, bool FULL, bool TAP>
__device__ __forceinline__ void tdp_group(float (&V)[P], uint32_t (&acc)[P], float &left, const uint32_t (&c)[P][8],
                                          int gi, uint32_t qs, const float (&mc)[P], int q, int lane, int x0, int y0,
                                          int nsteps, const EdgeLink &el, int &gc, float *tap0, float *tap1,
                                          long long *dst)
{
    (void)dst;
    const float4 q0 = lds128(qs + 32 * gi), q1 = lds128(qs + 32 * gi + 16);
    const float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    float v[P][8];
#pragma unroll
    for (int j = 0; j < P; ++j)
#pragma unroll
        for (int s = 0; s < 8; ++s)   // tts.py:495: y_square - y_mu_double + (mu_square + const)
            v[j][s] = __fadd_rn(__fadd_rn(qv[s], __uint_as_float(c[j][s])), mc[j]);
    if (TAP) {
#pragma unroll
        for (int s = 0; s < 8; ++s)
            if (8 * gi + s < nsteps) {
                if (tap0) tap0[8 * gi + s] = v[0][s];
                if (P > 1 && tap1) tap1[8 * gi + s] = v[P - 1][s];
            }
    }
    // my left neighbour's last token at these 8 frames (warp 0: a constant block of -1e9)
    const uint32_t slot = (uint32_t)gc & 15u;
    DPSTAT(long long c0 = clock64();)
    if (q > 0) mbar_wait_s(el.bar_in + slot * 8, (uint32_t)(gc >> 4) & 1u);
    DPSTAT(dst[1] += clock64() - c0;)
    const uint32_t ein = el.in + (q > 0 ? slot * 32 : 0);
    const float4 e0 = lds128(ein), e1 = lds128(ein + 16);
    const float ev[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
    uint32_t ga[P];
#pragma unroll
    for (int j = 0; j < P; ++j) ga[j] = 0u;
    float out[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        out[s] = 0.0f;
        if (FULL || 8 * gi + s < nsteps) {
            float nxt = 0.0f;
#pragma unroll
            for (int j = P - 1; j >= 0; --j) {
                const float up = (j == 0) ? left : V[j > 0 ? j - 1 : 0];   // V[x-1, y-1]
                dir_bit(up, V[j], ga[j], 1u << s);
                float nv = __fadd_rn(fmaxf(up, V[j]), v[j][s]);
                if (DIAG) nv = (x0 + j <= y0 + 8 * gi + s) ? nv : kNeg;   // x > y: not reachable yet
                V[j] = nv;
#ifdef MAS_EXP_NOSHFL   // timing experiment only (wrong results): the recurrence without its cross-lane hop
                if (j == P - 1) nxt = nv;
#else
                if (j == P - 1) nxt = __shfl_up_sync(kFull, nv, 1);        // in flight during the rest
#endif
            }
            left = (lane == 0) ? ev[s] : nxt;
            out[s] = V[P - 1];
        }
    }
#pragma unroll
    for (int j = 0; j < P; ++j) acc[j] |= ga[j] << (8 * gi);
    if (q < 3) {   // publish my last token (lane 31) for the warp on my right
        if (lane == 31) {
            sts128(el.out + slot * 32, out[0], out[1], out[2], out[3]);
            sts128(el.out + slot * 32 + 16, out[4], out[5], out[6], out[7]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_s(el.bar_out + slot * 8);
    }
    ++gc;
}

__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// One tile = 32 frames in four groups of 8: the accumulator columns of the next group are in
// flight (tcgen05.ld into the other register buffer) while a group is processed; the accumulator
// buffer is released as soon as its last columns are in registers.
template <int P, bool DIAG, bool FULL, bool TAP>
__device__ __forceinline__ void tdp_tile(float (&V)[P], uint32_t (&acc)[P], float &left, uint32_t dtm, uint32_t qs,
                                         const float (&mc)[P], int q, int lane, int x0, int y0, int nsteps,
                                         const EdgeLink &el, int &gc, uint64_t *d_empty_bar, float *tap0, float *tap1,
                                         long long *dst)
{
    uint32_t ca[P][8], cb[P][8];
    auto load = [&](uint32_t (&c)[P][8], int gi) {
#pragma unroll
        for (int j = 0; j < P; ++j) tmem_ld8(dtm + 32 * j + 8 * gi, c[j]);
    };
    auto release = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d_empty_bar);
    };
#define TDP_GROUP(c, gi) tdp_group<P, DIAG, FULL, TAP>(V, acc, left, c, gi, qs, mc, q, lane, x0, y0, nsteps, el, gc, tap0, tap1, dst)
    DPSTAT(long long c0 = clock64();)
    load(ca, 0);
    tmem_wait_ld();
    DPSTAT(dst[0] += clock64() - c0;)
    if (FULL) {
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {   // two groups per trip: the register buffers alternate without copies
            load(cb, 2 * h + 1);
            TDP_GROUP(ca, 2 * h);
            DPSTAT(c0 = clock64();)
            tmem_wait_ld();
            DPSTAT(dst[2] += clock64() - c0;)
            if (h == 0) load(ca, 2);
            else release();
            TDP_GROUP(cb, 2 * h + 1);
            if (h == 0) tmem_wait_ld();
        }
    } else {
        const int ng = (nsteps + 7) >> 3;
#pragma unroll 1
        for (int gi = 0; gi < ng; ++gi) {
            if (gi + 1 < ng) load(cb, gi + 1);
            else release();
            TDP_GROUP(ca, gi);
            if (gi + 1 < ng) {
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < P; ++j)
#pragma unroll
                    for (int s = 0; s < 8; ++s) ca[j][s] = cb[j][s];
            }
        }
    }
#undef TDP_GROUP
}

__global__ void __launch_bounds__(kTcThreads, 1) mas_prior_tc2_kernel(const PriorTcArgs a)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const TcLayout &L = a.lay;
    const int F = a.F, Fp = L.Fp, T_x = a.T_x, NB = L.nb;
    const int64_t T_y = a.T_y;
    float *slabs = reinterpret_cast<float *>(smem + L.off_slabs);
    float *staging = reinterpret_cast<float *>(smem + L.off_staging);   // [kTcStage][Fp][32]
    uint32_t *bits_a = reinterpret_cast<uint32_t *>(smem + L.off_bits);
    uint32_t *bits_b = bits_a + (size_t)L.nch * L.xrows;
    float *ysq = reinterpret_cast<float *>(smem + L.off_ysq);
    float *musq = reinterpret_cast<float *>(smem + L.off_musq);   // [kTcMsq][256]
    int *first = reinterpret_cast<int *>(smem + L.off_first);
    int *dur = reinterpret_cast<int *>(smem + L.off_dur);
    uint32_t *zbuf = reinterpret_cast<uint32_t *>(smem + L.off_zero);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.off_bars);
    // barrier map
    uint64_t *slab_full = bars + 8, *slab_free = bars + 12;
    uint64_t *d_full = bars + 16, *d_empty = bars + 20;
    uint64_t *a_ready = bars + 24, *a_free = bars + 30;   // [2] each: one pair per 128-token M tile
    volatile int *ctrl = reinterpret_cast<volatile int *>(bars + 32);   // [0] zdone [2] bt_done [4..7] fwd_done per DP warp
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bars + 40);
    uint64_t *edge_bar = bars + 64;                             // [3 links][16 groups]
    float *edge = reinterpret_cast<float *>(bars + 128);        // [3 links][16 groups][8 frames]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int part_floats = Fp * kTileY;              // one K-major part: [Fp/8 k steps][32 frames][8 features]
    const int slab_floats = 2 * part_floats;          // hi then lo
    const float cst = (float)(-0.5 * 1.8378770664093453 * (double)F);  // -0.5*log(2*pi)*F, tts.py:484

    if (tid == 0) {
        for (int s = 0; s < 4; ++s) {
            mbar_init(&slab_full[s], 1);
            mbar_init(&slab_free[s], 1);
            mbar_init(&d_full[s], 1);
            mbar_init(&d_empty[s], 4);      // the four DP warps
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_ready[i], 4);
            mbar_init(&a_free[i], 1);
        }
        for (int i = 0; i < 48; ++i) mbar_init(&edge_bar[i], 1);
        for (int i = 0; i < 8; ++i) ctrl[i] = 0;
        for (int i = 0; i < 8; ++i) edge[384 + i] = kNeg;
        mbar_fence_init();
    }
    for (int i = tid; i < kTcZeroBytes / 4; i += kTcThreads) zbuf[i] = 0u;
    for (int i = tid; i < kTcSlabs * slab_floats; i += kTcThreads) slabs[i] = 0.0f;
    for (int i = tid; i < kTcStage * Fp * kTileY; i += kTcThreads) staging[i] = 0.0f;  // rows F..Fp-1 stay zero
    fence_proxy_async_smem();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)),
                     "r"(kTcTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = __shfl_sync(kFull, *tslot, 0);

    auto geometry = [&](int u, int &tx, int &ty, int &ntiles, bool &degenerate) {
        tx = min(max(a.t_x[u], 0), T_x);
        ty = min(max(a.t_y[u], 0), a.T_y);
        degenerate = tx > ty && ty >= 1;
        const bool active = tx >= 1 && ty >= 1 && !degenerate;
        ntiles = active ? (ty + kTileY - 1) / kTileY : 0;
    };
    // tokens per TMEM lane of an utterance, and the M tiles it uses (token x = lane * P + M tile)
    auto tokens_per_lane = [&](int tx) -> int { return tx > 128 ? 2 : 1; };
    auto tile_mask = [&](int tx, int ty, int t) -> int {
        (void)ty;
        (void)t;
        return tokens_per_lane(tx) == 2 ? 3 : 1;
    };
    volatile int *zdone = ctrl, *bt_done = ctrl + 2, *fwd_done = ctrl + 4;   // fwd_done[4]: one per DP warp
    const bool son = a.stats != nullptr;
    long long *so = son ? a.stats + (size_t)blockIdx.x * 32 : nullptr;
    const int bslots = L.bits_slots;
    auto bits_of = [&](int k) -> uint32_t * { return ((k & 1) && bslots == 2) ? bits_b : bits_a; };

    if (warp < 4) {
        // ======================= DP warps (TMEM lane quarter q = warp) =======================
        const int q = warp;
        const int Lg = 32 * q + lane;                       // TMEM lane = global DP lane
        const uint32_t lane_base = tbase + ((uint32_t)(32 * q) << 16);
        EdgeLink el;   // [384..391] of `edge`: -1e9, the "left neighbour" of token 0
        el.in = smem_u32(edge + (q > 0 ? (q - 1) * 128 : 384));
        el.out = smem_u32(edge + (q < 3 ? q * 128 : 0));
        el.bar_in = smem_u32(edge_bar + (q > 0 ? (q - 1) * 16 : 0));
        el.bar_out = smem_u32(edge_bar + (q < 3 ? q * 16 : 0));
        int g = 0, gc = 0, k = 0, ka = 0;
        TcStat s_all(son && q == 0), s_w(son && q == 0), s_bt(son && q == 0);
        long long dpst[3] = {0, 0, 0};   // MAS_TC2_DPSTATS: first ld of a tile, edge waits, later lds
        s_all.begin();
        for (int u = blockIdx.x; u < a.B; u += gridDim.x, ++k) {
            int tx, ty, ntiles;
            bool degenerate;
            geometry(u, tx, ty, ntiles, degenerate);
            if (ntiles > 0) {
                const int P = tokens_per_lane(tx);
                const int x0 = Lg * P;
                s_bt.begin();
                while (*bt_done < k - bslots + 1) __nanosleep(32);
                s_bt.end();
                __threadfence_block();
                uint32_t *bits = bits_of(k);
                float V[2] = {kNeg, kNeg}, ms[2] = {0.0f, 0.0f};
                uint32_t acc[2] = {0u, 0u};
                float left = (Lg == 0) ? 0.0f : kNeg;       // frame 0: v_prev(x=0) = 0, everything else -1e9
                const float *msq = musq + (ka % kTcMsq) * 256;
                for (int t = 0; t < ntiles; ++t, ++g) {
                    const int b = g % NB;
                    s_w.begin();
                    mbar_wait(&d_full[b], (g / NB) & 1);
                    s_w.end();
                    tc_fence_after();
                    if (t == 0) {   // -0.5|mu|^2 + const of my tokens (visible: the movers' arrive precedes the MMAs)
                        ms[0] = (x0 < tx) ? __fadd_rn(msq[x0], cst) : 0.0f;
                        ms[1] = (P == 2 && x0 + 1 < tx) ? __fadd_rn(msq[x0 + 1], cst) : 0.0f;
                    }
                    const uint32_t dtm = lane_base + L.col_d + b * 64;
                    const uint32_t qs = smem_u32(ysq + (g % kTcYsq) * kTileY);
                    const int y0 = t * kTileY;
                    const int nsteps = min(kTileY, ty - y0);
                    const bool diag = y0 < tx;
                    float *tap0 = nullptr, *tap1 = nullptr;
                    if (a.lp_out) {
                        if (x0 < tx) tap0 = a.lp_out + ((int64_t)u * T_x + x0) * T_y + y0;
                        if (P == 2 && x0 + 1 < tx) tap1 = a.lp_out + ((int64_t)u * T_x + x0 + 1) * T_y + y0;
                    }
#define TDP_ARGS left, dtm, qs, M, q, lane, x0, y0, nsteps, el, gc, &d_empty[b], tap0, tap1, dpst
                    if (P == 2) {
                        float(&W)[2] = V;
                        uint32_t(&A)[2] = acc;
                        const float(&M)[2] = ms;
                        if (a.lp_out) tdp_tile<2, true, false, true>(W, A, TDP_ARGS);
                        else if (nsteps != kTileY) tdp_tile<2, true, false, false>(W, A, TDP_ARGS);
                        else if (diag) tdp_tile<2, true, true, false>(W, A, TDP_ARGS);
                        else tdp_tile<2, false, true, false>(W, A, TDP_ARGS);
                    } else {
                        float(&W)[1] = *reinterpret_cast<float(*)[1]>(&V[0]);
                        uint32_t(&A)[1] = *reinterpret_cast<uint32_t(*)[1]>(&acc[0]);
                        const float(&M)[1] = *reinterpret_cast<const float(*)[1]>(&ms[0]);
                        if (a.lp_out) tdp_tile<1, true, false, true>(W, A, TDP_ARGS);
                        else if (nsteps != kTileY) tdp_tile<1, true, false, false>(W, A, TDP_ARGS);
                        else if (diag) tdp_tile<1, true, true, false>(W, A, TDP_ARGS);
                        else tdp_tile<1, false, true, false>(W, A, TDP_ARGS);
                    }
#undef TDP_ARGS
                    // x == y always steps down (core.pyx:34 `index == y`), token 0 never does.
                    if (diag) {
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                            if (j < P && ((x0 + j) >> 5) == t) acc[j] |= 1u << ((x0 + j) & 31);
                    }
                    if (Lg == 0) acc[0] = 0u;
                    uint32_t *dst = bits + (size_t)t * L.xrows + Lg;
                    dst[0] = acc[0];
                    acc[0] = 0u;
                    if (P == 2) {
                        dst[128] = acc[1];
                        acc[1] = 0u;
                    }
                }
                // total alignment score: token tx-1 = lane (tx-1)/P, slot (tx-1)%P
                const int ql = (tx - 1) / P, qj = (tx - 1) - ql * P;
                if (Lg == ql && a.score) a.score[u] = V[qj];
                ++ka;
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) fwd_done[q] = k + 1;
        }
        s_all.end();
        if (son && lane == 0 && q == 0) { so[0] = s_all.acc; so[1] = s_w.acc; so[2] = g; so[3] = k; so[28] = s_bt.acc; }
        DPSTAT(if (son && lane == 0 && (q == 0 || q == 3)) { long long *d = so + (q == 0 ? 13 : 16); d[0] = dpst[0]; d[1] = dpst[1]; d[2] = dpst[2]; })
    } else if (warp == kTcBack) {
        // ======================= backtrack warp =======================
        int k = 0;
        TcStat b_bt(son), b_w(son), b_out(son);
        for (int u = blockIdx.x; u < a.B; u += gridDim.x, ++k) {
            int tx, ty, ntiles;
            bool degenerate;
            geometry(u, tx, ty, ntiles, degenerate);
            for (int x = lane; x < T_x; x += 32) dur[x] = 0;
            b_w.begin();
            while (fwd_done[0] <= k || fwd_done[1] <= k || fwd_done[2] <= k || fwd_done[3] <= k) __nanosleep(32);
            b_w.end();
            b_bt.begin();
            __threadfence_block();
            __syncwarp();
            if (ntiles > 0) {
#ifndef MAS_EXP_NOBACK   // (timing experiment only when defined: no path)
                if (lane == 0) backtrack_bits(bits_of(k), L.xrows, tx, ty, first, dur, true, 7);
#endif
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
            if (lane == 0) *bt_done = k + 1;
            b_bt.end();
            b_out.begin();
            if (a.path) {
                while (*zdone <= k) __nanosleep(64);
                __threadfence_block();
            }
            char *pb = a.path ? static_cast<char *>(a.path) + (int64_t)u * T_x * T_y * a.path_esize : nullptr;
            write_path_ones(pb, a.durations ? a.durations + (int64_t)u * T_x : nullptr, first, dur, T_x, T_y,
                            a.path_esize, a.one, lane, 32);
            write_frame_idx(a.frame_idx ? a.frame_idx + (int64_t)u * T_y : nullptr, first, dur, T_x, ty, a.T_y,
                            lane, 32);
            __syncwarp();
            b_out.end();
        }
        if (son && lane == 0) { so[29] = b_bt.acc; so[30] = b_w.acc; so[31] = b_out.acc; }
    } else if (warp == kTcLoader || warp == kTcLoader2) {
        // ======================= slab loader =======================
        // y[:, 32t..32t+31] -> staging buffer in shared memory (cp.async, 16-byte pieces, the natural
        // [feature][32 frames] layout) -> one pass per slab that transposes into the K-major hi/lo slab:
        // lane (r, c) = (lane / 8, lane % 8) owns frames 4c..4c+3 of the feature groups {4(r+4n)..+3}:
        // four 16-byte reads, a 4x4 transpose in registers, the tf32 split, four + four 16-byte stores
        // (frame rows), plus -0.5|y|^2 per frame.  Loops stay rolled on purpose: with six different
        // roles resident it is the instruction cache, not a pipe, that a warp waits for (ncu:
        // stall_no_inst dominated the unrolled version of this pass).
        // Also: the bulk (TMA) zero-fill of the dense output path.
        const bool vec16 = (T_y % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.y) & 15) == 0);
        const int c = lane & 7, r = lane >> 3;
        TcStat l_all(son), l_free(son), l_fin(son), l_cp(son), l_ld(son), l_fence(son), l_issue(son);
        l_all.begin();
        auto finish = [&](int gg) {
            l_cp.begin();
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            l_cp.end();
            l_fin.begin();
            __syncwarp();                 // every lane's copies of this slab have landed
            const float *st = staging + (size_t)(gg & 1) * Fp * kTileY + 4 * c;   // tile parity == owning loader
            float *hi = slabs + (size_t)(gg % kTcSlabs) * slab_floats, *lo = hi + part_floats;
            float q0 = 0.0f, q1 = 0.0f, q2 = 0.0f, q3 = 0.0f;   // frames 4c .. 4c+3
#ifdef MAS_EXP_NOLOADERPASS   // timing experiment only (wrong results)
            if (gg < 0)
#endif
#pragma unroll 1
            for (int fg = r; 4 * fg < Fp; fg += 4) {   // feature group: features 4fg .. 4fg+3
                const float *sp = st + (fg << 7);
                const float4 a0 = *reinterpret_cast<const float4 *>(sp);
                const float4 a1 = *reinterpret_cast<const float4 *>(sp + 32);
                const float4 a2 = *reinterpret_cast<const float4 *>(sp + 64);
                const float4 a3 = *reinterpret_cast<const float4 *>(sp + 96);
                q0 = __fmaf_rn(a0.x, a0.x, q0); q0 = __fmaf_rn(a1.x, a1.x, q0); q0 = __fmaf_rn(a2.x, a2.x, q0); q0 = __fmaf_rn(a3.x, a3.x, q0);
                q1 = __fmaf_rn(a0.y, a0.y, q1); q1 = __fmaf_rn(a1.y, a1.y, q1); q1 = __fmaf_rn(a2.y, a2.y, q1); q1 = __fmaf_rn(a3.y, a3.y, q1);
                q2 = __fmaf_rn(a0.z, a0.z, q2); q2 = __fmaf_rn(a1.z, a1.z, q2); q2 = __fmaf_rn(a2.z, a2.z, q2); q2 = __fmaf_rn(a3.z, a3.z, q2);
                q3 = __fmaf_rn(a0.w, a0.w, q3); q3 = __fmaf_rn(a1.w, a1.w, q3); q3 = __fmaf_rn(a2.w, a2.w, q3); q3 = __fmaf_rn(a3.w, a3.w, q3);
                const float rows[4][4] = {{a0.x, a1.x, a2.x, a3.x}, {a0.y, a1.y, a2.y, a3.y},
                                          {a0.z, a1.z, a2.z, a3.z}, {a0.w, a1.w, a2.w, a3.w}};
                // k step fg/2, 16-byte half fg%2 (swapped on rows with (j >> 2) & 1 == c & 1)
                const int base = ((fg >> 1) << 8) + (c << 5) + (((fg & 1) ^ (c & 1)) << 2);
#pragma unroll
                for (int e = 0; e < 4; ++e) {   // frame row j = 4c + e
                    const float4 h = make_float4(tf32_rn(rows[e][0]), tf32_rn(rows[e][1]), tf32_rn(rows[e][2]),
                                                 tf32_rn(rows[e][3]));
                    *reinterpret_cast<float4 *>(hi + base + (e << 3)) = h;
                    *reinterpret_cast<float4 *>(lo + base + (e << 3)) =
                        make_float4(tf32_rn(rows[e][0] - h.x), tf32_rn(rows[e][1] - h.y), tf32_rn(rows[e][2] - h.z),
                                    tf32_rn(rows[e][3] - h.w));
                }
            }
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {   // the four feature-group phases r of a frame
                q0 += __shfl_xor_sync(kFull, q0, o);
                q1 += __shfl_xor_sync(kFull, q1, o);
                q2 += __shfl_xor_sync(kFull, q2, o);
                q3 += __shfl_xor_sync(kFull, q3, o);
            }
            if (r == 0)   // tts.py:488-490  y_square
                *reinterpret_cast<float4 *>(ysq + (gg % kTcYsq) * kTileY + 4 * c) =
                    make_float4(-0.5f * q0, -0.5f * q1, -0.5f * q2, -0.5f * q3);
            l_fence.begin();
            fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&slab_full[gg % kTcSlabs]);
            l_fence.end();
            l_fin.end();
        };
        // Two loader warps split the tiles by parity of the CTA-lifetime tile index; each owns one
        // staging buffer: copy of its tile g in flight -> (at its next tile) transpose pass -> next copy.
        const int lp = (warp == kTcLoader) ? 0 : 1;
        int g = 0, pend = -1, k = 0;
        for (int u = blockIdx.x; u < a.B; u += gridDim.x, ++k) {
            int tx, ty, ntiles;
            bool degenerate;
            geometry(u, tx, ty, ntiles, degenerate);
            const float *yb = a.y + (int64_t)u * F * T_y;
            char *pb = a.path ? static_cast<char *>(a.path) + (int64_t)u * T_x * T_y * a.path_esize : nullptr;
            const int64_t pbytes = a.path ? (int64_t)T_x * T_y * a.path_esize : 0;
            const bool zbulk = bulk_zero_ok(pb, pbytes);
            if (lp == 0) {   // the even loader also clears the dense output path of the utterance
                if (zbulk) {
                    zero_fill_bulk_part(pb, pbytes, 0, 1, zbuf, kTcZeroBytes, lane, 32);
                    bulk_commit();
                } else {
                    zero_fill_part(pb, pbytes, 0, 1, lane, 32);
                }
            }
            for (int t = 0; t < ntiles; ++t, ++g) {
                if ((g & 1) != lp) continue;
                if (pend >= 0) finish(pend);   // my previous tile: its copy has had a whole trip to land
                const int s = g % kTcSlabs;
                l_free.begin();
                if (g >= kTcSlabs) mbar_wait_relaxed(&slab_free[s], ((g / kTcSlabs) - 1) & 1);  // MMAs of tile g-kTcSlabs done
                l_free.end();
                float *dst = staging + (size_t)lp * Fp * kTileY;
                const int y0 = t * kTileY;
                l_issue.begin();
                if (vec16) {
                    const int left = ty - (y0 + 4 * c);
                    const uint32_t bytes = left >= 4 ? 16u : (left > 0 ? 4u * left : 0u);
                    const float *src = yb + (bytes ? y0 + 4 * c : 0);
#pragma unroll 5
                    for (int f = r; f < F; f += 4) cp_async16(dst + (f << 5) + 4 * c, src + (int64_t)f * T_y, bytes);
                } else {
                    const int y = y0 + lane;
                    const uint32_t bytes = y < ty ? 4u : 0u;
                    const float *src = yb + (y < ty ? y : 0);
#pragma unroll 4
                    for (int f = 0; f < F; ++f) cp_async4(dst + (f << 5) + lane, src + (int64_t)f * T_y, bytes);
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                pend = g;
                l_issue.end();
            }
            if (lp == 0) {
                // the utterance's zeros must be in HBM before the backtrack warp writes its 1-cells; the
                // tile in flight is finished first so that the wait for the stores delays nobody
                if (pend >= 0) { finish(pend); pend = -1; }
                if (zbulk) bulk_wait_all();
                __threadfence_block();
                __syncwarp();
                if (lane == 0) *zdone = k + 1;
            }
        }
        if (pend >= 0) finish(pend);
        l_all.end();
        if (son && lane == 0 && lp == 0) {
            so[4] = l_all.acc; so[5] = l_free.acc; so[6] = l_fin.acc; so[7] = l_cp.acc;
            so[23] = l_ld.acc; so[24] = l_fence.acc; so[25] = l_issue.acc;
        }
    } else if (warp == kTcMma) {
        // ======================= MMA issuer =======================
        // The whole warp runs the loop (converged, warp-uniform operands); one elected lane issues.
        {
            // instruction descriptor: D fp32, A/B tf32, A and B K-major, N = 32, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
            const int ksteps = Fp >> 3;
            int g = 0, ka = 0;
            TcStat m_all(son), m_a(son), m_s(son), m_d(son), m_i(son);
            m_all.begin();
            for (int u = blockIdx.x; u < a.B; u += gridDim.x) {
                int tx, ty, ntiles;
                bool degenerate;
                geometry(u, tx, ty, ntiles, degenerate);
                if (ntiles == 0) continue;
                // A operand readiness / release is tracked per M tile: the first ~4 tiles of an utterance
                // only touch tokens < 128, and its last tiles only tokens >= 128, so the movers can
                // refill one half of TMEM while the other is still (or already) being multiplied
                bool a_ok[2] = {false, false}, a_rel[2] = {false, false};
                // last tile whose band still reaches into M tile 0 (tokens < 128)
                const int t0_last = ntiles - 1;   // tokens are interleaved over the M tiles: both live to the end
                for (int t = 0; t < ntiles;) {
                    // a pair of tiles (g, g+1): up to four independent accumulators in flight
                    // (accumulator c = 2*e + i: tile e of the pair, M tile i); everything the issue
                    // loop needs lives in registers -- the single issuing thread is the bottleneck
                    const int np = (t + 1 < ntiles) ? 2 : 1;
                    uint32_t dcol[4];
                    uint64_t bh[2], bl[2];
                    int am = 0;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        if (e >= np) break;
                        const int gg = g + e, s = gg % kTcSlabs, b = gg % NB;
                        m_s.begin();
                        mbar_wait(&slab_full[s], (gg / kTcSlabs) & 1);
                        m_s.end();
                        m_d.begin();
                        if (gg >= NB) mbar_wait(&d_empty[b], ((gg / NB) - 1) & 1);  // epilogue drained tile gg-NB
                        m_d.end();
                        const uint32_t sb = smem_u32(slabs + (size_t)s * slab_floats);
                        bh[e] = tc_bdesc(sb);
                        bl[e] = tc_bdesc(sb + (uint32_t)part_floats * 4u);
                        am |= tile_mask(tx, ty, t + e) << (2 * e);
                        dcol[2 * e] = tbase + L.col_d + (b * 2) * 32;
                        dcol[2 * e + 1] = dcol[2 * e] + 32;
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        if (!a_ok[i] && ((am >> i) & 1 || (am >> (i + 2)) & 1)) {
                            m_a.begin();
                            mbar_wait(&a_ready[i], ka & 1);   // this half of mu_x is in TMEM
                            m_a.end();
                            a_ok[i] = true;
                        }
                    tc_fence_after();
                    m_i.begin();
                    // Everything the issue loop uses is made provably warp-uniform first (mas_prior_tc.cu)
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) dcol[c4] = uniform_u32(dcol[c4]);
                    uint32_t bhl[2], bll[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        bhl[e] = uniform_u32((uint32_t)bh[e]);
                        bll[e] = uniform_u32((uint32_t)bl[e]);
                    }
                    const uint32_t bdhi = (uint32_t)(tc_bdesc(0) >> 32);
                    am = (int)uniform_u32((uint32_t)am);
                    const uint32_t ahi0 = uniform_u32(tbase + L.col_ahi), alo0 = uniform_u32(tbase + L.col_alo);
                    const uint32_t fpu = uniform_u32((uint32_t)Fp);
                    auto issue_all = [&](auto ks_tag, auto am_tag) {
                        constexpr int KS = decltype(ks_tag)::value;   // 0 = run-time k-step count (rolled loop)
                        constexpr int AM = decltype(am_tag)::value;   // 0 = run-time accumulator mask
                        const int nk = KS ? KS : ksteps;
                        const int m = AM ? AM : am;
#pragma unroll
                        for (int pass = kExpPass0; pass < 3; ++pass) {
                            const uint32_t abase = (pass == 0) ? alo0 : ahi0;
                            const uint32_t b0 = (pass == 1) ? bll[0] : bhl[0], b1 = (pass == 1) ? bll[1] : bhl[1];
#pragma unroll
                            for (int j = 0; j < nk; ++j) {
                                const uint32_t acc = ((pass - kExpPass0) | j) != 0;
                                const uint32_t ac = abase + 8 * j;
                                const uint32_t bo = (uint32_t)(64 * j);
                                if (m & 1) tc_mma_ts2(dcol[0], ac, b0 + bo, bdhi, idesc, acc);
                                if (m & 2) tc_mma_ts2(dcol[1], ac + fpu, b0 + bo, bdhi, idesc, acc);
                                if (m & 4) tc_mma_ts2(dcol[2], ac, b1 + bo, bdhi, idesc, acc);
                                if (m & 8) tc_mma_ts2(dcol[3], ac + fpu, b1 + bo, bdhi, idesc, acc);
                            }
                        }
                    };
                    // one elected lane issues the whole pair; with the k-step count known at compile time
                    // the loop unrolls and the operands of all 120 MMAs are immediates off a few registers
                    if (elect_one()) {
                        using K10 = std::integral_constant<int, 10>;
                        if (ksteps != 10) issue_all(std::integral_constant<int, 0>{}, std::integral_constant<int, 0>{});
                        else if (am == 15) issue_all(K10{}, std::integral_constant<int, 15>{});
                        else if (am == 5) issue_all(K10{}, std::integral_constant<int, 5>{});
                        else if (am == 3) issue_all(K10{}, std::integral_constant<int, 3>{});
                        else if (am == 1) issue_all(K10{}, std::integral_constant<int, 1>{});
                        else issue_all(K10{}, std::integral_constant<int, 0>{});
                    }
                    __syncwarp();
                    if (elect_one()) {
                        for (int e = 0; e < np; ++e) {
                            tc_commit(&slab_free[(g + e) % kTcSlabs]);   // slab reusable once these MMAs have read it
                            tc_commit(&d_full[(g + e) % NB]);            // accumulators ready for the epilogue warps
                        }
                    }
                    __syncwarp();
                    m_i.end();
                    t += np;
                    g += np;
                    if (!a_rel[0] && t > t0_last) {   // no later tile touches tokens < 128
                        if (!a_ok[0]) mbar_wait(&a_ready[0], ka & 1);
                        if (elect_one()) tc_commit(&a_free[0]);
                        __syncwarp();
                        a_rel[0] = a_ok[0] = true;
                    }
                }
                // every barrier sees every utterance, used or not
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (!a_ok[i]) mbar_wait(&a_ready[i], ka & 1);
                    if (!a_rel[i]) {
                        if (elect_one()) tc_commit(&a_free[i]);
                        __syncwarp();
                    }
                }
                ++ka;
            }
            m_all.end();
            if (son && lane == 0) { so[8] = m_all.acc; so[9] = m_a.acc; so[10] = m_s.acc; so[11] = m_d.acc; so[12] = m_i.acc; }
        }
    } else if (warp >= 8 && warp < 12) {
        // ======================= mu_x movers (TMEM lane quarter q = warp - 8) =======================
        const int q = warp - 8;
        const uint32_t lane_base = tbase + ((uint32_t)(32 * q) << 16);
        const bool aon = son && q == 0;
        TcStat a_all(aon), a_ld(aon), a_w(aon), a_st(aon);
        a_all.begin();
        int ka = 0;
        for (int u = blockIdx.x; u < a.B; u += gridDim.x) {
            int tx, ty, ntiles;
            bool degenerate;
            geometry(u, tx, ty, ntiles, degenerate);
            if (ntiles == 0) continue;
            const int P = tokens_per_lane(tx);
            const float *mub = a.mu_x + (int64_t)u * F * T_x;
            float *msq = musq + (ka % kTcMsq) * 256;
            for (int i = 0; i < 2; ++i) {
                const bool any = i < P && (32 * q) * P + i < tx;   // some of this warp's tokens exist
                const int x = (32 * q + lane) * P + i;         // token of (M tile i, TMEM lane 32q + lane)
                const bool xv = x < tx;
                // every feature of "my" token into registers: all loads in flight at once, issued
                // while the previous utterance is still being multiplied
                float v[96];
                a_ld.begin();
                if (any) {
#pragma unroll
                    for (int f = 0; f < 96; ++f)
                        v[f] = (xv && f < F) ? __ldg(mub + (int64_t)f * T_x + x) : 0.0f;
                }
                a_ld.end();
                a_w.begin();
                if (ka > 0) mbar_wait_relaxed(&a_free[i], (ka - 1) & 1, 128);  // previous utterance's MMAs have read this half
                a_w.end();
                tc_fence_after();
                a_st.begin();
                if (any) {
                    float s = 0.0f;
#pragma unroll
                    for (int f0 = 0; f0 < 96; f0 += 8) {
                        if (f0 >= Fp) break;
                        uint32_t rh[8], rl[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float m = v[f0 + e];
                            s = __fmaf_rn(m, m, s);
                            const float h = tf32_rn(m);
                            rh[e] = __float_as_uint(h);
                            rl[e] = __float_as_uint(tf32_rn(m - h));
                        }
                        tmem_st8(lane_base + L.col_ahi + i * Fp + f0, rh);
                        tmem_st8(lane_base + L.col_alo + i * Fp + f0, rl);
                    }
                    msq[x] = -0.5f * s;   // tts.py:494  mu_square
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_ready[i]);
                a_st.end();
            }
            ++ka;
        }
        a_all.end();
        if (aon && lane == 0) { so[19] = a_all.acc; so[20] = a_ld.acc; so[21] = a_w.acc; so[22] = a_st.acc; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(kTcTmemCols));
}

cudaError_t launch_from_prior_tc2(const PriorTcArgs &a, cudaStream_t st)
{
    void (*k)(const PriorTcArgs) = mas_prior_tc2_kernel;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.lay.total);
    if (e != cudaSuccess) return e;
    if (a.lp_out) {
        // parity tap: the kernel below writes every cell it produces (= what the DP consumed); the
        // cells it never forms (padding, empty and degenerate utterances) come from the plain kernel
        e = launch_log_prior(a.mu_x, a.y, a.lp_out, a.B, a.F, a.T_x, a.T_y, st);
        if (e != cudaSuccess) return e;
    }
    // persistent: one CTA per SM; `sm_limit` (mas_set_sm_limit) leaves SMs free for a concurrent kernel,
    // e.g. the NCCL all-gather of the previous step's durations
    const int grid = std::min(a.B, std::max(1, sm_count() - sm_reserve()));
    k<<<grid, kTcThreads, a.lay.total, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

}  // namespace mas
