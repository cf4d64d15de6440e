// mas_prior_tc.cu -- fused Gaussian log-prior + MAS with the prior on the 5th-gen tensor cores.
//
// Same contract as mas_prior_kernel (mas_prior.cu; replaces tts.py:483-505), different engine
// for the cross term sum_f mu[f,x]*y[f,j]: the reference computes it with an fp32 GEMM
// (torch.matmul, tts.py:491-493), and so does this kernel -- tcgen05.mma kind::tf32 with the
// 3xTF32 split, which restores fp32-level accuracy from tf32 products:
//
//     a = a_hi + a_lo,  b = b_hi + b_lo   (a_hi = tf32(a), a_lo = tf32(a - a_hi), same for b)
//     a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi          (dropped a_lo*b_lo ~ 2^-22 |a*b|)
//
// accumulated in fp32 in tensor memory (measured on B200, profiles/microbench/tc_prior.cu:
// max error 3e-5 on sums of magnitude 35, vs 1e-5 for a sequential fp32 FMA chain and 1e-2 for
// plain tf32).  The FMA pipe was the binding resource of the CUDA-core kernel (80 FMA per
// cell); here one utterance's mu_x lives in TENSOR MEMORY as the A operand (hi and lo parts,
// one column per feature, one lane per token) and each 32-frame slab of y is the B operand
// in shared memory, so a tile costs 60 MMAs of 128x32x8 and no CUDA-core arithmetic beyond
// the three adds of tts.py:495.
//
// Persistent CTA per SM, 14 warps (kTc* below), all hand-offs are mbarriers:
//   warps 0-3   epilogue (TMEM lane quarter = warp id): per tile tcgen05.ld the accumulator,
//               add the y / mu / const terms and write the swizzled tile the DP warps consume
//   warps 4, 13 DP warps (mas_dp.cuh dp_forward2: the token axis split over 64 lanes)
//   warp 5      MMA issuer (one elected lane): tiles are issued in PAIRS, interleaving the MMAs of
//               their (up to four) independent accumulators -- back-to-back MMAs into the same
//               accumulator serialise at ~100 cycles each (measured), independent ones overlap
//   warps 6, 7  slab loaders (alternate tiles): cp.async y slab -> staging -> one transposing pass
//               into the K-major SWIZZLE_32B hi/lo slab, -0.5|y|^2 per frame; the even loader also
//               issues the bulk (TMA) zero fill of the dense output path
//   warps 8-11  mu_x movers (TMEM lane quarter = warp id - 8): run one utterance AHEAD -- hold
//               the next utterance's mu_x in registers and, the moment its last MMA has retired,
//               split it into tf32 hi/lo and tcgen05.st it into TMEM; also produce -0.5|mu|^2
//   warp 12     backtrack warp (one utterance behind, second direction-bit buffer): path ones,
//               durations, frame index, peer-memory rows of the fused all-gather; clears the dense path of
//               the last of several utterances of a CTA while its tiles run
//
// Token axes of 257..512 (template parameter CS = 2 or 4): one thread-block CLUSTER per utterance, CTA h owns
// the tokens [h xs, (h+1) xs) with the same fourteen roles; the recurrence and the backtrack cross CTAs through
// distributed shared memory (mas_dp.cuh dp_forward_chain / backtrack_bits_window), the direction words live in
// the caller's workspace.  DESIGN.md 4.3 / 4.3c have the measurements (role cycles, per-tile timeline,
// what bounds the kernel) behind every choice below.
#include <algorithm>
#include <type_traits>

#include "mas_dp.cuh"
#include "mas_internal.h"

namespace mas {

namespace {

constexpr int kTcThreads = 448;
// warps 0-3 epilogue, 8-11 mu_x movers (TMEM lane quarter = warp % 4 is a hardware rule); the rest is
// placed by scheduler (warp % 4): the DP warp shares its scheduler with the latency-bound backtrack
// warp, not with a loader
constexpr int kTcDp = 4, kTcMma = 5, kTcLoader = 6, kTcLoader2 = 7, kTcBack = 12, kTcDp2 = 13;
constexpr int kTcSlabs = 4;       // y slabs (hi+lo) in flight
constexpr int kTcLag = 1;         // a slab is finished one loader iteration after its copy was issued
constexpr int kTcYsq = 8;         // ring of per-slab -0.5|y|^2 vectors (> kTcSlabs + D buffers)
constexpr int kTcStage = 2;        // staging buffers [F][32 frames] behind the cp.async copies: one per loader warp
constexpr int kTcMsq = 4;         // ring of per-utterance -0.5|mu|^2 vectors (> utterances an accumulator lags)
constexpr int kTcZeroBytes = 4096;
constexpr int kTcTmemCols = 512;
// cluster mode: inbound boundary ring [kXRing][32] floats (1 KB), then in_full[kXRing], out_empty[kXRing],
// bt_full[4] mbarriers and bt_in[4] ints (hand-over of the backtrack between CTAs); at +2048 the backtrack's
// window of direction words (32 tokens x kBtWinC chunks)
constexpr int kTcClBytes = 4096;

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
// The same with the shared-memory descriptor given as two 32-bit halves: the upper half is the same
// constant for every MMA of the kernel and the k step only moves the 14-bit address field of the
// lower one, so the issue loop adds 32-bit immediates instead of carrying 64-bit sums.
__device__ __forceinline__ void tc_mma_ts2(uint32_t d_tmem, uint32_t a_tmem, uint32_t bdesc_lo, uint32_t bdesc_hi,
                                           uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 bd, {%2, %3};\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t}" ::"r"(d_tmem),
                 "r"(a_tmem), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "r"(acc)
                 : "memory");
}
#ifdef MAS_TC_BF16CORR
// Experimental split (-DMAS_TC_BF16CORR): the two correction terms of the 3-term product in BF16 on kind::f16 MMAs
// (K = 16 per MMA: 5 + 5 MMAs instead of 10 + 10), the main term a_hi*b_hi stays tf32:
//     a*b ~= bf16(a - a_hi) * bf16(b)  +  bf16(a) * bf16(b - b_hi)  +  a_hi * b_hi
// 20 instead of 30 MMAs per accumulator chain, the same TMEM columns and slab bytes.
__device__ __forceinline__ void tc_mma_ts2_f16(uint32_t d_tmem, uint32_t a_tmem, uint32_t bdesc_lo, uint32_t bdesc_hi,
                                               uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 bd, {%2, %3};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, p;\n\t}" ::"r"(d_tmem),
                 "r"(a_tmem), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo_elem, float hi_elem)   // element with the LOWER k index in the low half
{
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}
#endif

// a value every lane holds identically, as a value the compiler KNOWS is warp-uniform (redux.sync
// writes a uniform register): arithmetic on it stays in the uniform datapath, where tcgen05.mma
// takes its operands -- otherwise each operand of each MMA is computed in vector registers and
// moved across (R2UR), ~8 instructions per MMA on the single issuing thread
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
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
TcLayout tc_layout(int F, int T_x, int T_y, int cluster)
{
    TcLayout L{};
    L.ok = 0;
    L.Fp = (F + 7) / 8 * 8;
    L.xrows = (T_x + 63) / 64 * 64;            // two DP warps: 64 lanes share the token axis
    L.nch = (T_y + 31) / 32;
    L.cluster = 1;
    L.xs = L.xrows;
    L.xr_tot = L.xrows;
    if (T_x > 512 || L.Fp > 96) return L;
#ifdef MAS_TC_BF16CORR
    if (L.Fp % 16) return L;
#endif
    if (T_x > 256) {
        // Long token axis: a thread-block cluster per utterance, CTA h owns the tokens [h xs, (h+1) xs) -- its own
        // M tiles of mu_x in its own tensor memory, its own tile ring and DP warps; the recurrence crosses the CTA
        // boundary through distributed shared memory (mas_dp.cuh dp_forward_chain).  Direction words go to the
        // workspace (L2): 4096 frames x 256 tokens of them would not fit next to the rings.
        L.cluster = (cluster == 2) ? 2 : 4;
        L.xs = 512 / L.cluster;
        L.xrows = L.xs;
    }
    const int mt = ((L.cluster > 1 ? L.xs : T_x) + 127) / 128;   // M tiles of 128 tokens per CTA
    L.col_ahi = 0;
    L.col_alo = mt * L.Fp;
    L.col_d = 2 * mt * L.Fp;
    L.nb = std::min(4, (kTcTmemCols - L.col_d) / 64);   // accumulator buffers of 2 x 32 columns
    if (L.nb < 2) return L;
    const size_t stage = (size_t)L.xrows * 128, bits = L.cluster > 1 ? 0 : (size_t)L.nch * L.xrows * 4;
    const size_t slabs = (size_t)kTcSlabs * 2 * L.Fp * 128;   // hi + lo, K-major: 1 KB per k step of 8 features
    const size_t misc = (size_t)kTcYsq * 128 + (size_t)kTcMsq * 256 * 4 + (((size_t)T_x * 8 + 15) & ~(size_t)15) + 16 +
                        kTcZeroBytes + 1024 + (size_t)kTcStage * L.Fp * 128 + (L.cluster > 1 ? kTcClBytes : 0);
    for (int slots = 2; slots >= 1 && !L.ok; --slots)
        for (int ns = 4; ns >= 2; --ns)
            if ((size_t)ns * stage + slots * bits + slabs + misc <= (size_t)kSmemBudget) {
                L.nstages = ns;
                L.bits_slots = slots;
                L.ok = 1;
                break;
            }
    if (!L.ok) return L;
    L.off_stages = 0;
    L.off_slabs = (size_t)L.nstages * stage;                  // both multiples of 1024
    L.off_staging = L.off_slabs + slabs;
    L.off_bits = L.off_staging + (size_t)kTcStage * L.Fp * 128;
    L.off_ysq = L.off_bits + (size_t)L.bits_slots * bits;
    L.off_musq = L.off_ysq + (size_t)kTcYsq * 128;
    L.off_first = L.off_musq + (size_t)kTcMsq * 256 * 4;
    L.off_dur = L.off_first + (size_t)T_x * 4;
    L.off_zero = (L.off_dur + (size_t)T_x * 4 + 15) & ~(size_t)15;
    L.off_bars = L.off_zero + kTcZeroBytes;
    L.off_cl = L.off_bars + 1024;
    L.total = L.off_cl + (L.cluster > 1 ? kTcClBytes : 0);
    // one CTA per SM: each CTA allocates all 512 TMEM columns
    L.total = std::max(L.total, (size_t)(kSmemBudget / 2 + 1024));
    return L;
}

// CS = CTAs per utterance: 1, or a thread-block cluster of 2 / 4 (launched with the cluster attribute) that
// splits the token axis (T_x > 256); everything cluster-specific is compiled out of the CS == 1 kernel.
template <int XPLMAX, int CS>
__global__ void __launch_bounds__(kTcThreads, 1) mas_prior_tc_kernel(const PriorTcArgs a)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr bool CL = CS > 1;
    const TcLayout &L = a.lay;
    const int crank = CL ? (int)cluster_ctarank() : 0;            // CTA h of the cluster
    const int cid = CL ? (int)blockIdx.x / CS : (int)blockIdx.x;  // persistent worker index
    const int ncl = CL ? (int)gridDim.x / CS : (int)gridDim.x;
    const int xs = CL ? L.xs : 0, xlo = crank * xs;               // this CTA's tokens: [xlo, xlo + xs)
    const int F = a.F, Fp = L.Fp, T_x = a.T_x, NS = L.nstages, NB = L.nb;
    const int64_t T_y = a.T_y;
    float *stages = reinterpret_cast<float *>(smem + L.off_stages);
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
    uint64_t *ring_full = bars, *ring_empty = bars + 4;
    uint64_t *slab_full = bars + 8, *slab_free = bars + 12;
    uint64_t *d_full = bars + 16, *d_empty = bars + 20;
    uint64_t *a_ready = bars + 24, *a_free = bars + 30;   // [2] each: one pair per 128-token M tile
    volatile int *ctrl = reinterpret_cast<volatile int *>(bars + 32);   // [0] zdone [1] fwd_done [2] bt_done
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bars + 40);
    uint64_t *edge_full = bars + 26;                            // [4] DP warp 0 -> DP warp 1, per tile
    float *edge = reinterpret_cast<float *>(bars + 64);         // [4 tiles][32 frames] boundary values

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int part_floats = Fp * kTileY;              // one K-major part: [Fp/8 k steps][32 frames][8 features]
    const int slab_floats = 2 * part_floats;          // hi then lo
    const float cst = (float)(-0.5 * 1.8378770664093453 * (double)F);  // -0.5*log(2*pi)*F, tts.py:484

    TileRing ring;
    ring.stages = stages;
    ring.full = ring_full;
    ring.empty = ring_empty;
    ring.nstages = NS;
    ring.stage_floats = L.xrows * kTileY;
    if (tid == 0) {
        for (int s = 0; s < 4; ++s) {
            mbar_init(&ring_full[s], 4);    // the four epilogue warps
            mbar_init(&ring_empty[s], 2);   // both DP warps
            mbar_init(&edge_full[s], 1);
            mbar_init(&slab_full[s], 1);
            mbar_init(&slab_free[s], 1);
            mbar_init(&d_full[s], 1);
            mbar_init(&d_empty[s], 4);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_ready[i], 4);
            mbar_init(&a_free[i], 1);
        }
        ctrl[0] = 0;
        ctrl[1] = 0;
        ctrl[2] = 0;
        ctrl[3] = 0;
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
#ifdef MAS_TC_TRACE
    if (blockIdx.x == 0 && threadIdx.x == 0) g_tc_trace = a.stats ? a.stats + 148 * 32 : nullptr;
    __syncthreads();
#endif

    // ntiles = tiles this CTA works on, the first of them global tile t_lo (frames 32 t_lo ..).  One CTA per
    // utterance: all of them.  Cluster: the tiles that hold band cells of the CTA's tokens (mas_dp.cuh).
    auto geometry = [&](int u, int &tx, int &ty, int &ntiles, bool &degenerate, int &t_lo) {
        tx = min(max(a.t_x[u], 0), T_x);
        ty = min(max(a.t_y[u], 0), a.T_y);
        degenerate = tx > ty && ty >= 1;
        const bool active = tx >= 1 && ty >= 1 && !degenerate;
        ntiles = active ? (ty + kTileY - 1) / kTileY : 0;
        t_lo = 0;
        if constexpr (CL) {
            if (ntiles > 0) {
                if (tx <= xlo) {
                    ntiles = 0;
                } else {
                    t_lo = xlo >> 5;
                    ntiles = min(ntiles - 1, (min(tx, xlo + xs) - 1 - tx + ty) >> 5) - t_lo + 1;
                }
            }
        }
    };
    // which of the CTA's 128-token M tiles hold band cells of (global) tile t (all of them when the parity tap is on)
    auto tile_mask = [&](int tx, int ty, int t) -> int {
        const int mt = CL ? (min(tx, xlo + xs) - xlo + 127) >> 7 : (tx + 127) >> 7;
        if (a.lp_out) return (1 << mt) - 1;
        const int lo = max(0, tx + t * kTileY - ty), hi = min(tx - 1, t * kTileY + kTileY - 1);
        int m = 0;
        for (int i = 0; i < mt; ++i)
            if (xlo + 128 * i <= hi && xlo + 128 * i + 127 >= lo) m |= 1 << i;
        return m;
    };
    // cluster mode: boundary ring and hand-over slots (kTcClBytes at off_cl)
    unsigned char *clb = smem + L.off_cl;
    uint64_t *x_in_full = reinterpret_cast<uint64_t *>(clb + 1024), *x_out_empty = x_in_full + kXRing;
    uint64_t *bt_full = x_out_empty + kXRing;
    int *bt_in = reinterpret_cast<int *>(bt_full + 4);
    if constexpr (CL) {
        if (tid == 0) {
            for (int s = 0; s < kXRing; ++s) {
                mbar_init(&x_in_full[s], 1);
                mbar_init(&x_out_empty[s], 1);
            }
            for (int s = 0; s < 4; ++s) mbar_init(&bt_full[s], 1);
            mbar_fence_init();
        }
        __syncwarp();
        cluster_sync_all();   // every CTA's barriers exist before anybody arrives on them remotely
    }
    volatile int *zdone = ctrl, *fwd_done = ctrl + 1, *bt_done = ctrl + 2, *fwd_done2 = ctrl + 3;
    const bool son = a.stats != nullptr;
    long long *so = son ? a.stats + (size_t)blockIdx.x * 32 : nullptr;
    const int bslots = L.bits_slots;
    auto bits_of = [&](int k) -> uint32_t * { return ((k & 1) && bslots == 2) ? bits_b : bits_a; };

    if (warp == kTcDp || warp == kTcDp2) {
        // ======================= DP warps: forward recurrence (mas_dp.cuh dp_forward2) =======================
        const int w = (warp == kTcDp) ? 0 : 1;
        int g = 0, k = 0;
        TcStat s_all(son);
        long long dp_wait = 0, ch_wait[2] = {0, 0};
        ChainCtx cx{};
        if constexpr (CL) {
            cx.in_ring = smem_u32(clb);
            cx.in_full = x_in_full;
            cx.out_empty = x_out_empty;
            cx.left_empty = map_to_cta(smem_u32(x_out_empty), (uint32_t)max(crank - 1, 0));
            cx.right_ring = map_to_cta(smem_u32(clb), (uint32_t)min(crank + 1, CS - 1));
            cx.right_full = map_to_cta(smem_u32(x_in_full), (uint32_t)min(crank + 1, CS - 1));
        }
        s_all.begin();
        for (int u = cid; u < a.B; u += ncl, ++k) {
            int tx, ty, ntiles, t_lo;
            bool degenerate;
            geometry(u, tx, ty, ntiles, degenerate, t_lo);
            if (ntiles > 0) {
                // cluster mode: the direction words live in the workspace (one region per utterance), the wait
                // only bounds how far the forward passes run ahead of the backtrack hand-overs (4 slots)
                while (*bt_done < k - bslots + 1) __nanosleep(32);
                __threadfence_block();
                int owns = 0;
                float score;
                if constexpr (CL) {
                    const int nt = (ty + kTileY - 1) / kTileY;
                    // warp 0: last tile the left CTA publishes a boundary for; warp 1: first tile the right CTA wants
                    const int in_last = (w == 0 && crank > 0) ? min(nt - 1, (xlo - 1 - tx + ty) >> 5) : t_lo - 2;
                    const int out_first = (w == 1 && crank < CS - 1 && tx > xlo + xs) ? ((xlo + xs) >> 5) - 1 : (1 << 30);
                    uint32_t *gbits = a.bits_ws + (size_t)u * L.nch * L.xr_tot + xlo;
                    score = chain_forward_dispatch<XPLMAX, true>(ring, gbits, L.xr_tot, tx, ty, min(tx, xlo + xs) - xlo,
                                                                 lane, w, g, edge, edge_full, xlo, t_lo, ntiles, in_last,
                                                                 out_first, cx, &owns, son ? ch_wait : nullptr);
                } else {
                    score = prior_forward2_dispatch<XPLMAX, true>(ring, bits_of(k), L.xrows, tx, ty, lane, w, g,
                                                                  edge, edge_full, &owns,
                                                                  son ? &dp_wait : nullptr);
                }
                g += ntiles;
                if (owns && lane == 0 && a.score) a.score[u] = score;
            }
            if constexpr (CL) __threadfence();   // direction words in the workspace: read back through L2 (ld.cg)
            else __threadfence_block();
            __syncwarp();
            if (lane == 0) *(w == 0 ? fwd_done : fwd_done2) = k + 1;
        }
        s_all.end();
        if (son && lane == 0 && w == 0) { so[0] = s_all.acc; so[1] = CL ? ch_wait[0] : dp_wait; so[2] = g; so[3] = k; so[14] = ch_wait[1]; }
        if (son && lane == 0 && w == 1) { so[26] = s_all.acc; so[27] = CL ? ch_wait[0] : dp_wait; so[15] = ch_wait[1]; }
    } else if (warp == kTcBack) {
        // ======================= backtrack warp =======================
        int k = 0, n_bt_in = 0, n_bt_out = 0;
        TcStat b_fwd(son), b_hand(son), b_walk(son), b_out(son);
        const uint32_t left_bt_in = CL ? map_to_cta(smem_u32(bt_in), (uint32_t)max(crank - 1, 0)) : 0u;
        const uint32_t left_bt_full = CL ? map_to_cta(smem_u32(bt_full), (uint32_t)max(crank - 1, 0)) : 0u;
        // cluster mode: this CTA's rows of the outputs
        const int x_beg = CL ? min(xlo, T_x) : 0, x_end = CL ? min(xlo + xs, T_x) : T_x;
        for (int u = cid; u < a.B; u += ncl, ++k) {
            int tx, ty, ntiles, t_lo;
            bool degenerate;
            geometry(u, tx, ty, ntiles, degenerate, t_lo);
            char *pb = a.path ? static_cast<char *>(a.path) + (int64_t)u * T_x * T_y * a.path_esize : nullptr;
            if constexpr (CL) {
                // Cluster mode: THIS warp clears the CTA's rows of the dense path (bulk stores), while the forward
                // pass of the utterance runs -- it has nothing else to do until then, and an utterance of 512 x 4096
                // is 8 MB of zeros: issued by a loader it would stall the tile pipeline for ~40 tiles.
                if (pb && x_end > x_beg && !a.path_zeroed) {
                    char *zb = pb + (int64_t)x_beg * T_y * a.path_esize;
                    const int64_t zbytes = (int64_t)(x_end - x_beg) * T_y * a.path_esize;
                    if (bulk_zero_ok(zb, zbytes)) {
                        zero_fill_bulk_part(zb, zbytes, 0, 1, zbuf, kTcZeroBytes, lane, 32);
                        bulk_commit();
                    } else {
                        zero_fill_part(zb, zbytes, 0, 1, lane, 32);
                    }
                }
            }
            const bool fill_here = !CL && pb && !a.path_zeroed && u + ncl >= a.B && k > 0;   // see the even loader
            if (fill_here) {
                const int64_t zbytes = (int64_t)T_x * T_y * a.path_esize;
                if (bulk_zero_ok(pb, zbytes)) {
                    zero_fill_bulk_part(pb, zbytes, 0, 1, zbuf, kTcZeroBytes, lane, 32);
                    bulk_commit();
                } else {
                    zero_fill_part(pb, zbytes, 0, 1, lane, 32);
                }
            }
            for (int x = lane; x < T_x; x += 32) dur[x] = 0;
            b_fwd.begin();
            while (*fwd_done <= k || *fwd_done2 <= k) __nanosleep(32);
            b_fwd.end();
            __threadfence_block();
            __syncwarp();
            if constexpr (CL) {
                if (ntiles > 0) {
                    // the CTA that owns token t_x - 1 starts; every other one takes over where its right
                    // neighbour left its tokens (frame handed over through distributed shared memory)
                    int y = ty - 1;
                    if (tx > xlo + xs) {
                        const int slot = n_bt_in & 3;
                        b_hand.begin();
                        mbar_wait_cluster(&bt_full[slot], (uint32_t)(n_bt_in >> 2) & 1u, 64);
                        b_hand.end();
                        y = bt_in[slot];
                        MAS_CHECK(y >= xlo + xs - 1 && y < ty);   // token xlo + xs - 1 ends at frame y >= its own index
                        ++n_bt_in;
                    }
                    const uint32_t *gbits = a.bits_ws + (size_t)u * L.nch * L.xr_tot + xlo;
                    b_walk.begin();
                    const int yl = backtrack_bits_window(gbits, L.xr_tot, min(tx, xlo + xs) - xlo, xlo,
                                                         min(tx, xlo + xs) - 1, y, first, dur,
                                                         reinterpret_cast<uint32_t *>(clb + 2048), lane);
                    b_walk.end();
                    if (crank > 0 && lane == 0) {
                        const int slot = n_bt_out & 3;
                        st_cluster_u32(left_bt_in + 4u * slot, (uint32_t)yl);
                        mbar_arrive_cluster(left_bt_full + 8u * slot);
                    }
                    if (crank > 0) ++n_bt_out;
                } else if (degenerate) {
                    // t_x > t_y (rare): every CTA walks the whole utterance and keeps its own rows
                    if (lane == 0) {
                        const float *mub = a.mu_x + (int64_t)u * F * T_x;
                        const float *yb = a.y + (int64_t)u * F * T_y;
                        auto val = [&](int x, int y) { return lp_cell(mub, yb, F, T_x, T_y, x, y, cst); };
                        backtrack_degenerate(val, tx, ty, first, dur);
                        if (a.score && crank == 0) a.score[u] = val(tx - 1, ty - 1);
                    }
                } else if (lane == 0 && a.score && crank == 0 && !(tx >= 1 && ty >= 1)) {
                    a.score[u] = 0.0f;
                }
            } else if (ntiles > 0) {
                if (lane == 0) backtrack_bits(bits_of(k), L.xrows, tx, ty, first, dur, true, 6);
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
            if constexpr (CL) {
                b_out.begin();
                bulk_wait_all();          // my zeros are in memory before my 1-cells
                __syncwarp();
            } else if (fill_here) {
                bulk_wait_all();          // my zeros are in memory before my 1-cells
                __syncwarp();
            } else if (a.path) {
                while (*zdone <= k) __nanosleep(64);
                __threadfence_block();
            }
            if constexpr (CL) {
                // rows [x_beg, x_end) only
                for (int x = x_beg + lane; x < x_end; x += 32) {
                    const int d = dur[x];
                    if (a.durations) a.durations[(int64_t)u * T_x + x] = d;
                    if (pb && d > 0) {
                        const int64_t e0 = (int64_t)x * T_y + first[x];
                        for (int kk = 0; kk < d; ++kk) st_one(pb, e0 + kk, a.path_esize, a.one);
                    }
                }
                if (a.frame_idx) {
                    int32_t *fi = a.frame_idx + (int64_t)u * T_y;
                    if (crank == 0)
                        for (int y = ty + lane; y < a.T_y; y += 32) fi[y] = -1;
                    for (int x = x_beg + lane; x < x_end; x += 32) {
                        const int d = dur[x], f0 = first[x];
                        for (int kk = 0; kk < d; ++kk) fi[f0 + kk] = x;
                    }
                }
                for (int p = 0; p < a.npeer; ++p) {
                    int32_t *row = a.peer[p] + (a.peer_row0 + u) * a.peer_stride;
                    for (int x = x_beg + lane; x < x_end; x += 32) row[x] = dur[x];
                    if (a.peer_fi[p]) {
                        int32_t *fi = a.peer_fi[p] + (a.peer_row0 + u) * a.peer_fi_stride;
                        if (crank == 0)
                            for (int y = ty + lane; y < a.T_y; y += 32) fi[y] = -1;
                        for (int x = x_beg + lane; x < x_end; x += 32) {
                            const int d = dur[x], f0 = first[x];
                            for (int kk = 0; kk < d; ++kk) fi[f0 + kk] = x;
                        }
                    }
                }
                __syncwarp();
                b_out.end();
                continue;
            }
            write_path_ones(pb, a.durations ? a.durations + (int64_t)u * T_x : nullptr, first, dur, T_x, T_y,
                            a.path_esize, a.one, lane, 32);
            write_frame_idx(a.frame_idx ? a.frame_idx + (int64_t)u * T_y : nullptr, first, dur, T_x, ty, a.T_y,
                            lane, 32);
            // fused all-gather (mas_peer_gather): the rows of this utterance go straight into row
            // row0 + u of every rank's buffers over NVLink peer memory (fire-and-forget stores)
            for (int p = 0; p < a.npeer; ++p) {
                MAS_CHECK(a.peer_stride >= T_x && a.peer_row0 >= 0);
                int32_t *row = a.peer[p] + (a.peer_row0 + u) * a.peer_stride;
                for (int x = lane; x < T_x; x += 32) row[x] = dur[x];
                if (a.peer_fi[p])
                    write_frame_idx(a.peer_fi[p] + (a.peer_row0 + u) * a.peer_fi_stride, first, dur, T_x, ty, a.T_y,
                                    lane, 32);
            }
            __syncwarp();
        }
        if (son && CL && lane == 0) { so[28] = b_hand.acc; so[29] = b_walk.acc; so[30] = b_fwd.acc; so[31] = b_out.acc; }
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
            // the slab is only needed now (the copy went to the staging buffer): waiting for it here instead of
            // before the copy was issued puts the copy's latency under the wait
            l_free.begin();
            if (gg >= kTcSlabs) mbar_wait_relaxed(&slab_free[gg % kTcSlabs], ((gg / kTcSlabs) - 1) & 1);  // MMAs of tile gg-kTcSlabs done
            l_free.end();
            l_fin.begin();
            __syncwarp();                 // every lane's copies of this slab have landed
            const float *st = staging + (size_t)(gg & 1) * Fp * kTileY + 4 * c;   // tile parity == owning loader
            float *hi = slabs + (size_t)(gg % kTcSlabs) * slab_floats, *lo = hi + part_floats;
            float q0 = 0.0f, q1 = 0.0f, q2 = 0.0f, q3 = 0.0f;   // frames 4c .. 4c+3
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
#ifdef MAS_TC_BF16CORR
                    // bf16 parts, K-major rows of 16 features (32 bytes), SWIZZLE_32B: k16 step fg/4, this lane's four
                    // features are bytes (fg%4)*8 .. +7 of the row; part 1 = bf16(y), part 2 = bf16(y - y_hi)
                    {
                        char *p16 = reinterpret_cast<char *>(lo) + ((fg >> 2) << 10) + ((4 * c + e) << 5) +
                                    (((((fg & 3) >> 1) ^ (c & 1))) << 4) + ((fg & 1) << 3);
                        *reinterpret_cast<uint2 *>(p16) = make_uint2(pack_bf16(rows[e][0], rows[e][1]), pack_bf16(rows[e][2], rows[e][3]));
                        *reinterpret_cast<uint2 *>(p16 + part_floats * 2) =
                            make_uint2(pack_bf16(rows[e][0] - h.x, rows[e][1] - h.y), pack_bf16(rows[e][2] - h.z, rows[e][3] - h.w));
                    }
#else
                    *reinterpret_cast<float4 *>(lo + base + (e << 3)) =
                        make_float4(tf32_rn(rows[e][0] - h.x), tf32_rn(rows[e][1] - h.y), tf32_rn(rows[e][2] - h.z),
                                    tf32_rn(rows[e][3] - h.w));
#endif
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
            MAS_TRACE(1, gg);
            l_fence.end();
            l_fin.end();
        };
        // Two loader warps split the tiles by parity of the CTA-lifetime tile index; each owns one
        // staging buffer: copy of its tile g in flight -> (at its next tile) transpose pass -> next copy.
        const int lp = (warp == kTcLoader) ? 0 : 1;
        int g = 0, pend = -1, k = 0;
        for (int u = cid; u < a.B; u += ncl, ++k) {
            int tx, ty, ntiles, t_lo;
            bool degenerate;
            geometry(u, tx, ty, ntiles, degenerate, t_lo);
            const float *yb = a.y + (int64_t)u * F * T_y;
            char *pb = a.path ? static_cast<char *>(a.path) + (int64_t)u * T_x * T_y * a.path_esize : nullptr;
            const int64_t pbytes = a.path ? (int64_t)T_x * T_y * a.path_esize : 0;
            const bool zbulk = bulk_zero_ok(pb, pbytes);
            for (int t = 0; t < ntiles; ++t, ++g) {
                if ((g & 1) != lp) continue;
                if (pend >= 0) finish(pend);   // my previous tile: its copy has had a whole trip to land
                float *dst = staging + (size_t)lp * Fp * kTileY;
                const int y0 = (t_lo + t) * kTileY;
                l_issue.begin();
                if (vec16) {
                    const int left = ty - (y0 + 4 * c);
                    const uint32_t bytes = left >= 4 ? 16u : (left > 0 ? 4u * left : 0u);
                    // 32-bit element offsets off a pinned base and a pinned stride (ptxas otherwise re-reads the
                    // kernel parameters and carries 64-bit products per copy: ~25 cycles of issue per cp.async)
                    const float *yp = yb;
                    uint32_t stride = 4u * (uint32_t)T_y;
                    asm volatile("" : "+l"(yp), "+r"(stride));
                    uint32_t off = (uint32_t)r * (uint32_t)T_y + (bytes ? (uint32_t)(y0 + 4 * c) : 0u);
                    float *d = dst + (r << 5) + 4 * c;
#pragma unroll 5
                    for (int f = r; f < F; f += 4) {
                        cp_async16(d, yp + off, bytes);
                        off += stride;
                        d += 128;
                    }
                } else {
                    const int y = y0 + lane;
                    const uint32_t bytes = y < ty ? 4u : 0u;
                    const float *src = yb + (y < ty ? y : 0);
#pragma unroll 4
                    for (int f = 0; f < F; ++f) cp_async4(dst + (f << 5) + lane, src + (int64_t)f * T_y, bytes);
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                MAS_TRACE(0, g);
                pend = g;
                l_issue.end();
            }
            if (!CL && lp == 0) {
                // The even loader also clears the dense output path of the utterance: a burst of bulk (TMA)
                // stores that holds it for ~25 k cycles.  It is issued AFTER the utterance's tiles: the
                // pipeline then has ~9 tiles of this utterance buffered downstream to work through while the
                // next utterance's even tiles wait, the CTA's first utterance starts without the bubble, and
                // for its last one the burst runs under the tail of the forward pass.  (With one utterance
                // per CTA -- batch-sharded over 8 GPUs -- the burst used to be fully exposed: 12.7 us of 69.)
                // The zeros must be in HBM before the backtrack warp writes its 1-cells (`zdone`); the tile in
                // flight is finished first so that the burst delays nobody who could run.
                if (pend >= 0) { finish(pend); pend = -1; }
                if (a.path_zeroed) {
                    // MAS_FLAG_PATH_ZEROED: the caller cleared the dense path (copy engine, under the previous step)
                } else if (u + ncl >= a.B && k > 0) {
                    // the LAST of several utterances of this CTA: nothing follows that could hide the burst, so the
                    // backtrack warp -- idle until the forward pass is over -- clears this path while the tiles run
                    // (batch-sharded steps with 4 utterances per CTA: 0.0432 -> 0.0419 ms per 128-utterance step).
                    // A CTA with a single utterance keeps the burst at the end: next to the pipeline fill of its
                    // only forward pass it costs more than it saves (one launch of 128: 0.085 -> 0.090 ms)
                } else if (zbulk) {
                    zero_fill_bulk_part(pb, pbytes, 0, 1, zbuf, kTcZeroBytes, lane, 32);
                    bulk_commit();
                    bulk_wait_all();
                } else {
                    zero_fill_part(pb, pbytes, 0, 1, lane, 32);
                }
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
            for (int u = cid; u < a.B; u += ncl) {
                int tx, ty, ntiles, t_lo;
                bool degenerate;
                geometry(u, tx, ty, ntiles, degenerate, t_lo);
                if (ntiles == 0) continue;
                // A operand readiness / release is tracked per M tile: the first ~4 tiles of an utterance
                // only touch tokens < 128, and its last tiles only tokens >= 128, so the movers can
                // refill one half of TMEM while the other is still (or already) being multiplied
                bool a_ok[2] = {false, false}, a_rel[2] = {false, false};
                // last tile whose band still reaches into M tile 0 (tokens < 128)
                const int t0_last = a.lp_out ? ntiles - 1 : min(ntiles - 1, ((xlo + 127 - tx + ty) >> 5) - t_lo);   // tap: every tile
#ifdef MAS_TC_BF16CORR
                if (false) {   // the staggered order is not written for the bf16 split
#else
                if (a.mma_stagger) {
#endif
                    // STAGGERED issue (profiles/r2_tc_tile_timeline.txt).  Issuing the tiles in pairs made tile g+2
                    // wait for the epilogue to drain tile g, i.e. for the whole pair (g, g+1) to finish in the
                    // tensor pipe: with three accumulator buffers issue and execution never overlapped (7.8 k cycles
                    // per pair, the pipe busy 5.8 k of them).  Here step s interleaves the SECOND half of tile s-1's
                    // MMA chain with the FIRST half of tile s's: still four independent accumulators in flight, but
                    // the tiles finish one at a time, half a step apart, and the buffer tile s needs (tile s-3's) was
                    // drained a whole step ago.  The chains themselves are unchanged (same terms, same order).
                    const int nacc = 3 * ksteps, half = nacc >> 1;     // MMAs per accumulator, first-half length
                    const uint32_t bdhi = (uint32_t)(tc_bdesc(0) >> 32);
                    const uint32_t ahi0 = uniform_u32(tbase + L.col_ahi), alo0 = uniform_u32(tbase + L.col_alo);
                    const uint32_t fpu = uniform_u32((uint32_t)Fp);
                    uint32_t dX[2] = {0u, 0u}, bXh = 0u, bXl = 0u;    // tile s-1: accumulators, slab descriptors (hi / lo part)
                    int amX = 0;
                    for (int s = 0; s <= ntiles; ++s) {
                        const bool hasY = s < ntiles;
                        const int gY = g + s, gX = g + s - 1;
                        uint32_t dY[2] = {0u, 0u}, bYh = 0u, bYl = 0u;
                        int amY = 0;
                        if (hasY) {
                            const int sl = gY % kTcSlabs, b = gY % NB;
                            m_s.begin();
                            mbar_wait(&slab_full[sl], (gY / kTcSlabs) & 1);
                            m_s.end();
                            m_d.begin();
                            if (gY >= NB) mbar_wait(&d_empty[b], ((gY / NB) - 1) & 1);  // epilogue drained tile gY-NB
                            m_d.end();
                            const uint32_t sbm = smem_u32(slabs + (size_t)sl * slab_floats);
                            bYh = uniform_u32((uint32_t)tc_bdesc(sbm));
                            bYl = uniform_u32((uint32_t)tc_bdesc(sbm + (uint32_t)part_floats * 4u));
                            amY = (int)uniform_u32((uint32_t)tile_mask(tx, ty, t_lo + s));
                            dY[0] = uniform_u32(tbase + L.col_d + (b * 2) * 32);
                            dY[1] = uniform_u32(tbase + L.col_d + (b * 2) * 32 + 32);
#pragma unroll
                            for (int i = 0; i < 2; ++i)
                                if (!a_ok[i] && ((amY >> i) & 1)) {
                                    m_a.begin();
                                    mbar_wait(&a_ready[i], ka & 1);   // this half of mu_x is in TMEM
                                    m_a.end();
                                    a_ok[i] = true;
                                }
                        }
                        tc_fence_after();
                        MAS_TRACE(2, gY);
                        m_i.begin();
                        // MMA number n (0 .. 3 ksteps - 1) of an accumulator's chain: pass n / ksteps (0: a_lo b_hi,
                        // 1: a_hi b_lo, 2: a_hi b_hi), k step n % ksteps
                        auto issue = [&](auto ks_tag, auto mx_tag, auto my_tag) {
                            constexpr int KS = decltype(ks_tag)::value;       // 0 = run-time k-step count
                            constexpr int MX = decltype(mx_tag)::value, MY = decltype(my_tag)::value;   // -1 = run-time mask
                            const int nk = KS ? KS : ksteps;
                            const int mx = MX >= 0 ? MX : amX, my = MY >= 0 ? MY : amY;
                            const int n1 = 3 * nk, h = n1 >> 1;
#pragma unroll
                            for (int j = 0; j < (KS ? (3 * KS + 1) / 2 : 1 << 30); ++j) {
                                if (!KS && j >= n1 - h) break;
                                if (j < h && my) {            // first half of tile s
                                    const int pass = j / nk, k = j - pass * nk;
                                    const uint32_t ab = ((pass == 0) ? alo0 : ahi0) + 8 * k;
                                    const uint32_t bb = ((pass == 1) ? bYl : bYh) + (uint32_t)(64 * k);
                                    if (my & 1) tc_mma_ts2(dY[0], ab, bb, bdhi, idesc, (uint32_t)(j != 0));
                                    if (my & 2) tc_mma_ts2(dY[1], ab + fpu, bb, bdhi, idesc, (uint32_t)(j != 0));
                                }
                                if (mx) {                     // second half of tile s-1
                                    const int n = h + j;
                                    const int pass = n / nk, k = n - pass * nk;
                                    const uint32_t ab = ((pass == 0) ? alo0 : ahi0) + 8 * k;
                                    const uint32_t bb = ((pass == 1) ? bXl : bXh) + (uint32_t)(64 * k);
                                    if (mx & 1) tc_mma_ts2(dX[0], ab, bb, bdhi, idesc, 1u);
                                    if (mx & 2) tc_mma_ts2(dX[1], ab + fpu, bb, bdhi, idesc, 1u);
                                }
                            }
                        };
                        if (elect_one()) {
                            using K10 = std::integral_constant<int, 10>;
                            using R = std::integral_constant<int, -1>;
                            if (ksteps != 10) issue(std::integral_constant<int, 0>{}, R{}, R{});
                            else if (amX == 3 && amY == 3) issue(K10{}, std::integral_constant<int, 3>{}, std::integral_constant<int, 3>{});
                            else if (amX == 1 && amY == 1) issue(K10{}, std::integral_constant<int, 1>{}, std::integral_constant<int, 1>{});
                            else if (amX == 2 && amY == 2) issue(K10{}, std::integral_constant<int, 2>{}, std::integral_constant<int, 2>{});
                            else issue(K10{}, R{}, R{});
                        }
                        __syncwarp();
                        if (s >= 1 && elect_one()) {
                            tc_commit(&slab_free[gX % kTcSlabs]);   // slab reusable once these MMAs have read it
                            tc_commit(&d_full[gX % NB]);            // tile s-1 complete: ready for the epilogue warps
                        }
                        __syncwarp();
                        MAS_TRACE(3, gX < 0 ? 0 : gX);
                        m_i.end();
                        if (!a_rel[0] && s >= 1 && s - 1 >= t0_last) {   // no later tile touches tokens < 128
                            if (!a_ok[0]) mbar_wait(&a_ready[0], ka & 1);
                            if (elect_one()) tc_commit(&a_free[0]);
                            __syncwarp();
                            a_rel[0] = a_ok[0] = true;
                        }
                        dX[0] = dY[0]; dX[1] = dY[1]; bXh = bYh; bXl = bYl; amX = amY;
                    }
                    g += ntiles;
                } else
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
                        am |= tile_mask(tx, ty, t_lo + t + e) << (2 * e);
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
                    MAS_TRACE(2, g);
                    if (np == 2) MAS_TRACE(2, g + 1);
                    m_i.begin();
                    // Everything the issue loop uses is made provably warp-uniform first (uniform_u32): the
                    // operands of all 120 MMAs are then uniform-register immediates off a few registers.
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) dcol[c4] = uniform_u32(dcol[c4]);
                    uint32_t bhl[2], bll[2];   // lower descriptor halves (hi / lo part of the slab)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        bhl[e] = uniform_u32((uint32_t)bh[e]);
                        bll[e] = uniform_u32((uint32_t)bl[e]);
                    }
                    const uint32_t bdhi = (uint32_t)(tc_bdesc(0) >> 32);
                    am = (int)uniform_u32((uint32_t)am);
                    const uint32_t ahi0 = uniform_u32(tbase + L.col_ahi), alo0 = uniform_u32(tbase + L.col_alo);
                    const uint32_t fpu = uniform_u32((uint32_t)Fp);
                    // small terms first: a_lo*b_hi + a_hi*b_lo, then a_hi*b_hi.  One k step = 8 features
                    // = 8 TMEM columns of A = 1 KB of B (64 descriptor address units).
                    // AM = the accumulators of the pair that exist (bit 2e + i: tile e, M tile i), as a
                    // compile-time mask for the common cases: predicating every MMA on a run-time mask
                    // costs the issuing thread more instructions than the MMA itself (0 = run-time mask)
                    auto issue_all = [&](auto ks_tag, auto am_tag) {
                        constexpr int KS = decltype(ks_tag)::value;   // 0 = run-time k-step count (rolled loop)
                        constexpr int AM = decltype(am_tag)::value;
                        const int nk = KS ? KS : ksteps;
                        const int m = AM ? AM : am;
#ifdef MAS_TC_BF16CORR
                        // pass 0: bf16(a_lo) * bf16(b), pass 1: bf16(a) * bf16(b_lo)  (kind::f16, nk/2 steps of 16 features),
                        // pass 2: a_hi * b_hi (kind::tf32, nk steps of 8).  A columns: lo16 / a16 parts hold Fp/2 columns per
                        // M tile; B: the bf16 parts follow the tf32 hi part in the slab (1 KB per step either way)
                        const uint32_t idesc16 = (1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
                        const uint32_t h2 = fpu >> 1;                       // Fp / 2 columns per M tile
                        const uint32_t a16base = alo0 + h2 * (uint32_t)((L.col_d - L.col_alo) / Fp);   // behind the lo16 parts of all M tiles
                        const uint32_t pb16 = (uint32_t)(part_floats * 2) >> 4;   // descriptor units from part 1 to part 2
#pragma unroll
                        for (int pass = 0; pass < 2; ++pass) {
                            const uint32_t abase = (pass == 0) ? alo0 : a16base;
                            const uint32_t b0 = bll[0] + (pass ? pb16 : 0u), b1 = bll[1] + (pass ? pb16 : 0u);
#pragma unroll
                            for (int j = 0; j < (nk >> 1); ++j) {
                                const uint32_t acc = (pass | j) != 0;
                                const uint32_t ac = abase + 8 * j;
                                const uint32_t bo = (uint32_t)(64 * j);
                                if (m & 1) tc_mma_ts2_f16(dcol[0], ac, b0 + bo, bdhi, idesc16, acc);
                                if (m & 2) tc_mma_ts2_f16(dcol[1], ac + h2, b0 + bo, bdhi, idesc16, acc);
                                if (m & 4) tc_mma_ts2_f16(dcol[2], ac, b1 + bo, bdhi, idesc16, acc);
                                if (m & 8) tc_mma_ts2_f16(dcol[3], ac + h2, b1 + bo, bdhi, idesc16, acc);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < nk; ++j) {
                            const uint32_t ac = ahi0 + 8 * j;
                            const uint32_t bo = (uint32_t)(64 * j);
                            if (m & 1) tc_mma_ts2(dcol[0], ac, bhl[0] + bo, bdhi, idesc, 1u);
                            if (m & 2) tc_mma_ts2(dcol[1], ac + fpu, bhl[0] + bo, bdhi, idesc, 1u);
                            if (m & 4) tc_mma_ts2(dcol[2], ac, bhl[1] + bo, bdhi, idesc, 1u);
                            if (m & 8) tc_mma_ts2(dcol[3], ac + fpu, bhl[1] + bo, bdhi, idesc, 1u);
                        }
#else
#pragma unroll
                        for (int pass = 0; pass < 3; ++pass) {
                            const uint32_t abase = (pass == 0) ? alo0 : ahi0;
                            const uint32_t b0 = (pass == 1) ? bll[0] : bhl[0], b1 = (pass == 1) ? bll[1] : bhl[1];
#pragma unroll
                            for (int j = 0; j < nk; ++j) {
                                const uint32_t acc = (pass | j) != 0;
                                const uint32_t ac = abase + 8 * j;
                                const uint32_t bo = (uint32_t)(64 * j);
                                if (m & 1) tc_mma_ts2(dcol[0], ac, b0 + bo, bdhi, idesc, acc);
                                if (m & 2) tc_mma_ts2(dcol[1], ac + fpu, b0 + bo, bdhi, idesc, acc);
                                if (m & 4) tc_mma_ts2(dcol[2], ac, b1 + bo, bdhi, idesc, acc);
                                if (m & 8) tc_mma_ts2(dcol[3], ac + fpu, b1 + bo, bdhi, idesc, acc);
                            }
                        }
#endif
                    };
                    // one elected lane issues the whole pair; with the k-step count known at compile time
                    // the loop unrolls and the operands of all 120 MMAs are immediates off a few registers
                    if (elect_one()) {
                        using K10 = std::integral_constant<int, 10>;
                        if (ksteps != 10) issue_all(std::integral_constant<int, 0>{}, std::integral_constant<int, 0>{});
                        else if (am == 15) issue_all(K10{}, std::integral_constant<int, 15>{});
                        else if (am == 5) issue_all(K10{}, std::integral_constant<int, 5>{});
                        else if (am == 10) issue_all(K10{}, std::integral_constant<int, 10>{});
                        else if (am == 3) issue_all(K10{}, std::integral_constant<int, 3>{});
                        else if (am == 1) issue_all(K10{}, std::integral_constant<int, 1>{});
                        else if (am == 2) issue_all(K10{}, std::integral_constant<int, 2>{});
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
                    MAS_TRACE(3, g);
                    if (np == 2) MAS_TRACE(3, g + 1);
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
        for (int u = cid; u < a.B; u += ncl) {
            int tx, ty, ntiles, t_lo;
            bool degenerate;
            geometry(u, tx, ty, ntiles, degenerate, t_lo);
            if (ntiles == 0) continue;
            const int mt = CL ? (min(tx, xlo + xs) - xlo + 127) >> 7 : (tx + 127) >> 7;   // this CTA's M tiles
            const float *mub = a.mu_x + (int64_t)u * F * T_x;
            float *msq = musq + (ka % kTcMsq) * 256;
            for (int i = 0; i < 2; ++i) {
                const bool any = i < mt && xlo + 128 * i + 32 * q < tx;   // some of this warp's tokens exist
                const int x = xlo + 128 * i + 32 * q + lane;
                const bool xv = x < tx;
                // the features of "my" token go through registers in two halves of 48 (all loads of a half in
                // flight at once; the first half is issued while the previous utterance is still being multiplied,
                // the second one is prefetched into L2 meanwhile).  Holding all 96 at once put the kernel over its
                // 128-register budget: the spilled values serialised the loads (15 k cycles to issue them).
                float s = 0.0f;
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    const int fb = 48 * h;
                    float v[48];
                    a_ld.begin();
                    if (any && fb < Fp) {
                        // running pointer + a stride pinned in a register: with the address formed from the kernel
                        // parameters per load, ptxas re-read them from the constant bank before every LDG (two LDC
                        // + their scoreboard wait per load: 80 cycles per load, 15 k cycles to issue them all)
                        // (32-bit element offsets off a uniform base: one IMAD.WIDE per address, no carry chains)
                        uint32_t stride = (uint32_t)T_x;
                        const float *mb = mub;
                        asm volatile("" : "+r"(stride), "+l"(mb));
                        uint32_t off = (uint32_t)fb * stride + (uint32_t)x;
                        const int nf = xv ? min(48, F - fb) : 0;
#pragma unroll
                        for (int f = 0; f < 48; ++f) {
                            v[f] = (f < nf) ? __ldg(mb + off) : 0.0f;
                            off += stride;
                        }
                        if (h == 0 && xv) {
#pragma unroll 4
                            for (int f = 48; f < F; ++f) {
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(mb + off));
                                off += stride;
                            }
                        }
                    }
                    a_ld.end();
                    if (h == 0) {
                        a_w.begin();
                        if (ka > 0) mbar_wait_relaxed(&a_free[i], (ka - 1) & 1, 128);  // previous utterance's MMAs have read this half
                        a_w.end();
                        tc_fence_after();
                    }
                    a_st.begin();
                    if (any && fb < Fp) {
#ifdef MAS_TC_BF16CORR
                        const int mtl = (L.col_d - L.col_alo) / Fp;        // M tiles of the layout
#pragma unroll
                        for (int f0 = 0; f0 < 48; f0 += 16) {
                            if (fb + f0 >= Fp) break;
                            uint32_t rh[16], rl[8], ra[8];
#pragma unroll
                            for (int e = 0; e < 16; ++e) {
                                const float m = v[f0 + e];
                                s = __fmaf_rn(m, m, s);
                                rh[e] = __float_as_uint(tf32_rn(m));
                            }
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const float m0 = v[f0 + 2 * e], m1 = v[f0 + 2 * e + 1];
                                rl[e] = pack_bf16(m0 - __uint_as_float(rh[2 * e]), m1 - __uint_as_float(rh[2 * e + 1]));
                                ra[e] = pack_bf16(m0, m1);
                            }
                            uint32_t r0[8], r1[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) { r0[e] = rh[e]; r1[e] = rh[8 + e]; }
                            tmem_st8(lane_base + L.col_ahi + i * Fp + fb + f0, r0);
                            tmem_st8(lane_base + L.col_ahi + i * Fp + fb + f0 + 8, r1);
                            tmem_st8(lane_base + L.col_alo + i * (Fp >> 1) + ((fb + f0) >> 1), rl);
                            tmem_st8(lane_base + L.col_alo + mtl * (Fp >> 1) + i * (Fp >> 1) + ((fb + f0) >> 1), ra);
                        }
#else
#pragma unroll
                        for (int f0 = 0; f0 < 48; f0 += 8) {
                            if (fb + f0 >= Fp) break;
                            uint32_t rh[8], rl[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const float m = v[f0 + e];
                                s = __fmaf_rn(m, m, s);
                                const float hh = tf32_rn(m);
                                rh[e] = __float_as_uint(hh);
                                rl[e] = __float_as_uint(tf32_rn(m - hh));
                            }
                            tmem_st8(lane_base + L.col_ahi + i * Fp + fb + f0, rh);
                            tmem_st8(lane_base + L.col_alo + i * Fp + fb + f0, rl);
                        }
#endif
                    }
                    a_st.end();
                }
                a_st.begin();
                if (any) msq[x - xlo] = -0.5f * s;   // tts.py:494  mu_square
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
    } else {
        // ======================= epilogue warps (TMEM lane quarter q = warp) =======================
        const int q = warp;
        const uint32_t lane_base = tbase + ((uint32_t)(32 * q) << 16);
        int g = 0, ka = 0;
        const bool eon = son && q == 0;
        TcStat e_all(eon), e_df(eon), e_re(eon), e_w(eon);
        e_all.begin();
        for (int u = cid; u < a.B; u += ncl) {
            int tx, ty, ntiles, t_lo;
            bool degenerate;
            geometry(u, tx, ty, ntiles, degenerate, t_lo);
            if (ntiles == 0) continue;
            const RowMap rm(CL ? min(tx, xlo + xs) - xlo : tx, 6);
            const float *msq = musq + (ka % kTcMsq) * 256;
            for (int t = 0; t < ntiles; ++t, ++g) {
                const int b = g % NB, sidx = g % NS;
                const int mask = tile_mask(tx, ty, t_lo + t);
                e_df.begin();
                mbar_wait_relaxed(&d_full[b], (g / NB) & 1, 32);
                if (q == 0) MAS_TRACE(4, g);
                e_df.end();
                tc_fence_after();
                e_w.begin();
                uint32_t acc[2][32];
                bool have[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    have[i] = (mask >> i & 1) && (xlo + 128 * i + 32 * q < tx);
                    if (have[i]) tmem_ld32(lane_base + L.col_d + (b * 2 + i) * 32, acc[i]);
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // -0.5|y|^2 of this tile is folded in BEFORE the accumulator is handed back.  Once d_empty is
                // released the MMA issuer -- and with it the loaders, kTcSlabs tiles further -- may run ahead
                // while this warp still waits for room in the DP ring below; the ysq ring (kTcYsq entries)
                // only covers the tiles up to that hand-back.  Reading it after the wait was a latent race:
                // the MAS_CHECK build, whose slower epilogue waits longer, returned a wrong prior in
                // test_fused_vs_oracle_configs[F=16] (profiles/r2_debug_checks.txt).
                {
                    const float *qs = ysq + (g % kTcYsq) * kTileY;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        if (!have[i]) continue;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float4 qq = *reinterpret_cast<const float4 *>(qs + 4 * c);
                            acc[i][4 * c + 0] = __float_as_uint(qq.x + __uint_as_float(acc[i][4 * c + 0]));
                            acc[i][4 * c + 1] = __float_as_uint(qq.y + __uint_as_float(acc[i][4 * c + 1]));
                            acc[i][4 * c + 2] = __float_as_uint(qq.z + __uint_as_float(acc[i][4 * c + 2]));
                            acc[i][4 * c + 3] = __float_as_uint(qq.w + __uint_as_float(acc[i][4 * c + 3]));
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_empty[b]);       // accumulator buffer free again
                e_w.end();
                e_re.begin();
                if (g >= NS) mbar_wait_relaxed(&ring_empty[sidx], ((g / NS) - 1) & 1, 32);  // DP consumed tile g-NS
                if (q == 0) MAS_TRACE(5, g);
                e_re.end();
                e_w.begin();
                float *tile = stages + (size_t)sidx * ring.stage_floats;
                const int y0 = (t_lo + t) * kTileY;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (!have[i]) continue;
                    const int x = xlo + 128 * i + 32 * q + lane;
                    if (x >= tx) continue;
                    const float ms = msq[x - xlo] + cst;   // -0.5|mu|^2 - 0.5 F log 2pi: one add per cell instead of two
                    const int pr = rm.row(x - xlo);
                    MAS_CHECK(pr >= 0 && pr < L.xrows && sidx >= 0 && sidx < NS);
                    float *row = tile + (pr << 5);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        float4 o;
                        // tts.py:495: ((y_square - y_mu_double) + mu_square) + const, the last two terms added first
                        // (the sum of 80 products in the accumulator is not the reference's rounding either; the
                        // parity bar of the fused path is 1e-5 relative, this is one ulp)
                        o.x = __uint_as_float(acc[i][4 * c + 0]) + ms;
                        o.y = __uint_as_float(acc[i][4 * c + 1]) + ms;
                        o.z = __uint_as_float(acc[i][4 * c + 2]) + ms;
                        o.w = __uint_as_float(acc[i][4 * c + 3]) + ms;
                        *reinterpret_cast<float4 *>(row + ((c ^ (pr & 7)) << 2)) = o;
                    }
                }
                if (a.lp_out) {   // parity tap (cold path): copy this warp's rows of the tile to HBM
                    __syncwarp();
                    for (int i = 0; i < 2; ++i) {
                        const int x = xlo + 128 * i + 32 * q + lane;
                        if (!have[i] || x >= tx) continue;
                        const int pr = rm.row(x - xlo);
                        const float *row = tile + (pr << 5);
                        float *tap = a.lp_out + ((int64_t)u * T_x + x) * T_y + y0;
                        for (int e = 0; e < kTileY && y0 + e < ty; ++e) tap[e] = row[(((e >> 2) ^ (pr & 7)) << 2) + (e & 3)];
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ring_full[sidx]);   // this warp's rows of tile g are written
                if (q == 0) MAS_TRACE(6, g);
                e_w.end();
            }
            ++ka;
        }
        e_all.end();
        if (eon && lane == 0) { so[13] = e_all.acc; so[16] = e_df.acc; so[17] = e_re.acc; so[18] = e_w.acc; }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CL) cluster_sync_all();   // nobody leaves while a neighbour may still arrive on its barriers
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(kTcTmemCols));
}

cudaError_t launch_from_prior_tc(const PriorTcArgs &a, cudaStream_t st)
{
    const int cs = a.lay.cluster;
    const int xplmax = ((cs > 1 ? a.lay.xs : a.T_x) + 63) / 64;   // tokens per lane of the two DP warps
    void (*k)(const PriorTcArgs) = nullptr;
    if (cs == 4) k = mas_prior_tc_kernel<2, 4>;
    else if (cs == 2) k = mas_prior_tc_kernel<4, 2>;
    else if (xplmax <= 2) k = mas_prior_tc_kernel<2, 1>;
    else if (xplmax <= 3) k = mas_prior_tc_kernel<3, 1>;
    else k = mas_prior_tc_kernel<4, 1>;
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
    // `utt_per_cta` (MAS_FLAG_UTT_PER_CTA): a caller that keeps several launches in flight (batch-sharded
    // steps on several streams) trades latency for throughput -- fewer CTAs, each running >= that many
    // utterances back to back, so that the pipeline fill/drain, the zero-fill burst and the backtrack of one
    // utterance overlap the forward pass of the next as they do in the one-launch-fills-the-chip case
    int grid = std::min(a.B, std::max(1, sm_count() - sm_reserve()));
    // never more CTAs than SMs: measured at B=1024, 256..512 CTAs of 2..4 utterances cost 12-30 % (0.37-0.44 ms
    // against 0.333) -- CTA start-up and the loss of the lock-step between the SMs' write bursts
    if (a.utt_per_cta > 1) grid = std::min(grid, std::max(1, (a.B + a.utt_per_cta - 1) / a.utt_per_cta));
    if (cs > 1) {
        // one cluster per utterance in flight: as many clusters as the device can hold at once (the GPC
        // structure decides: not every SM can be part of a 4-cluster), each running its utterances back to back
        cudaLaunchConfig_t cfg{};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.blockDim = dim3(kTcThreads);
        cfg.dynamicSmemBytes = a.lay.total;
        cfg.stream = st;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cfg.gridDim = dim3((unsigned)cs);
        int ncl = 0;
        e = cudaOccupancyMaxActiveClusters(&ncl, k, &cfg);
        if (e != cudaSuccess) return e;
        if (ncl < 1) return cudaErrorLaunchOutOfResources;
        ncl = std::min(ncl, std::max(1, (sm_count() - sm_reserve()) / cs));
        ncl = std::min(ncl, a.B);
        if (a.utt_per_cta > 1) ncl = std::min(ncl, std::max(1, (a.B + a.utt_per_cta - 1) / a.utt_per_cta));
        cfg.gridDim = dim3((unsigned)(ncl * cs));
        e = cudaLaunchKernelEx(&cfg, k, a);
        count_launch();
        return e != cudaSuccess ? e : cudaGetLastError();
    }
    k<<<grid, kTcThreads, a.lay.total, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

}  // namespace mas
