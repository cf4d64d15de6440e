// tc_prior.cu -- de-risking microbenchmark for the tensor-core log-prior (tcgen05, kind::tf32).
//
// Computes D[m, n] = sum_k mu[k][m] * y[k][n]  (M = 256 tokens as two M=128 tiles, N = 32 frames,
// K = 80 features) three ways and compares each with an fp64 reference:
//   v0  3xTF32, all operands from shared memory (SS):  A_hi*B_hi + A_lo*B_hi + A_hi*B_lo
//   v1  3xTF32, A_lo read from TENSOR MEMORY (TS) -- halves the shared memory mu_x needs
//   v2  plain TF32 (A_hi*B_hi only) -- shows what the split buys
// Operands are MN-major (the layout mu_x [F,T_x] / y [F,T_y] already have in HBM) in the
// SWIZZLE_128B canonical layout: panels of [K rows][32 elements] with 128-byte rows, 16-byte
// chunk c of row k stored at chunk c ^ (k & 7).  Also times a long MMA stream.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_prior tc_prior.cu ; run: ./tc_prior
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

constexpr int K = 80, M = 256, N = 32;
constexpr int PANEL = K * 128;  // bytes of one [K][32] fp32 panel

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 1)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
    d |= (uint64_t)layout << 61;  // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
    return d;
}

__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ float tf32_rn(float x)
{
    uint32_t b = __float_as_uint(x);
    b = (b + 0x1000u) & 0xffffe000u;
    return __uint_as_float(b);
}
__device__ __forceinline__ int sw_index(int k, int n)  // float index inside a [K][32] panel
{
    // SWIZZLE_128B_BASE32B (the only MN-major layout tf32 operands may use): rows of 128 B,
    // 32-byte unit u of row k stored at unit u ^ (k & 3)
    return k * 32 + ((((n >> 3) ^ (k & 3))) << 3) + (n & 7);
}

__global__ void __launch_bounds__(192) tc_test(const float *mu, const float *y, float *out, long long *cyc, int reps)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    float *A_hi = reinterpret_cast<float *>(smem);
    float *A_lo = reinterpret_cast<float *>(smem + 8 * PANEL);
    float *B_hi = reinterpret_cast<float *>(smem + 16 * PANEL);
    float *B_lo = reinterpret_cast<float *>(smem + 17 * PANEL);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 18 * PANEL);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(smem + 18 * PANEL + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < K * M; i += blockDim.x) {
        const int k = i / M, m = i % M;
        const float v = mu[i], h = tf32_rn(v);
        const int p = m >> 5;
        A_hi[p * (PANEL / 4) + sw_index(k, m & 31)] = h;
        if (m < 128) {  // K-major copy of tile 0: [kblock][m][32 k] swizzled rows of 128 B
            const int kb = k >> 5, kk = k & 31;
            A_lo[kb * 4096 + m * 32 + ((((kk >> 2) ^ (m & 7))) << 2) + (kk & 3)] = h;
        }
    }
    for (int i = tid; i < 16 * 128 * 32 / 32; i += blockDim.x) {}
    for (int i = tid; i < 128 * 16; i += blockDim.x) {  // zero pad k = 80..95 of block 2
        const int m = i >> 4, kk = 16 + (i & 15);
        A_lo[2 * 4096 + m * 32 + ((((kk >> 2) ^ (m & 7))) << 2) + (kk & 3)] = 0.0f;
    }
    for (int i = tid; i < K * N; i += blockDim.x) {
        const int k = i / N, n = i % N;
        const float v = y[i], h = tf32_rn(v);
        B_hi[sw_index(k, n)] = h;
        B_lo[sw_index(k, n)] = tf32_rn(v - h);
        {
            const int kb = k >> 5, kk = k & 31;
            A_lo[3 * 4096 + kb * 1024 + n * 32 + ((((kk >> 2) ^ (n & 7))) << 2) + (kk & 3)] = h;
        }
    }
    for (int i = tid; i < 32 * 16; i += blockDim.x) {
        const int n = i >> 4, kk = 16 + (i & 15);
        A_lo[3 * 4096 + 2 * 1024 + n * 32 + ((((kk >> 2) ^ (n & 7))) << 2) + (kk & 3)] = 0.0f;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 1)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 2)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = *tslot;
    if (tid == 0) cyc[2] = tbase;
    // TMEM map (columns): D v0: [0,64)  D v1: [64,128)  D v2: [128,192)  A_lo tile0: [256,336) tile1: [336,416)
    if (warp < 4) {
        for (int t = 0; t < 2; ++t) {
            const int m = 128 * t + 32 * warp + lane;
            for (int k0 = 0; k0 < K; k0 += 16) {
                uint32_t r[16];
                for (int q = 0; q < 16; ++q) {
                    const float v = mu[(k0 + q) * M + m];
                    r[q] = __float_as_uint(tf32_rn(v - tf32_rn(v)));
                }
                const uint32_t ta = tbase + ((uint32_t)(32 * warp) << 16) + 256 + 80 * t + k0;
                asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                             ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                               "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
            }
        }
        for (int c0 = 0; c0 < 192; c0 += 16) {
            const uint32_t sv = __float_as_uint(123.0f);
            const uint32_t ta = tbase + ((uint32_t)(32 * warp) << 16) + c0;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
                         ::"r"(ta), "r"(sv) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const uint32_t idesc_ss = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc_ts = idesc_ss & ~(1u << 15);  // A from TMEM is K-major by construction
    auto issue = [&](int variant, uint32_t dcol) {
        for (int t = 0; t < 2; ++t) {
            const uint32_t d = tbase + dcol + 32 * t;
            for (int j = 0; j < K / 8; ++j)
                mma_ss(d, make_desc(smem_u32(A_hi) + 4 * t * PANEL + j * 1024, PANEL, 512),
                       make_desc(smem_u32(B_hi) + j * 1024, 1024, 512), idesc_ss, j > 0);
            if (variant == 2) continue;
            for (int j = 0; j < K / 8; ++j) {
                if (variant == 0)
                    mma_ss(d, make_desc(smem_u32(A_lo) + 4 * t * PANEL + j * 1024, PANEL, 512),
                           make_desc(smem_u32(B_hi) + j * 1024, 1024, 512), idesc_ss, 1);
                else
                    mma_ts(d, tbase + 256 + 80 * t + 8 * j, make_desc(smem_u32(B_hi) + j * 1024, 1024, 512), idesc_ts, 1);
            }
            for (int j = 0; j < K / 8; ++j)
                mma_ss(d, make_desc(smem_u32(A_hi) + 4 * t * PANEL + j * 1024, PANEL, 512),
                       make_desc(smem_u32(B_lo) + j * 1024, 1024, 512), idesc_ss, 1);
        }
    };
    if (warp == 4 && lane == 0) {
        issue(1, 0);   // MN-major 3xTF32 with A_lo from TMEM, both tiles -> cols [0,64)
        issue(2, 128); // MN-major hi*hi only -> cols [128,192)
        {              // K-major hi*hi, tile 0 -> cols [64,96)
            const uint32_t idesc_k = idesc_ss & ~((1u << 15) | (1u << 16));
            for (int j = 0; j < 10; ++j) {
                const uint32_t kb = j >> 2, ko = (j & 3) * 32;
                mma_ss(tbase + 64, make_desc(smem_u32(A_lo) + kb * 16384 + ko, 16, 1024, 2),
                       make_desc(smem_u32(A_lo) + 3 * 16384 + kb * 4096 + ko, 16, 1024, 2), idesc_k, j > 0);
            }
        }
        commit(bar);
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) {
        for (int v = 0; v < 3; ++v)
            for (int t = 0; t < 2; ++t) {
                uint32_t r[32];
                const uint32_t ta = tbase + ((uint32_t)(32 * warp) << 16) + 64 * v + 32 * t;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                             "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                               "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                               "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                               "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                             : "r"(ta));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int m = 128 * t + 32 * warp + lane;
                for (int n = 0; n < N; ++n) out[(v * M + m) * N + n] = __uint_as_float(r[n]);
            }
    }
    // ---- throughput: `reps` x (one tile's 60 MMAs of the TS variant), one commit at the end
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 4 && lane == 0) {
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) issue(1, 64);
        const long long t1 = clock64();
        commit(bar + 1);
        mbar_wait(bar + 1, 0);
        const long long t2 = clock64();
        cyc[0] = t1 - t0;
        cyc[1] = t2 - t0;
    }
    // ---- cost per MMA vs N (TS and SS), D at columns [0, N), operands: whatever is in smem/TMEM
    if (warp == 4 && lane == 0) {
        int slot = 3;
        uint32_t ph = 1;
        for (int mode = 0; mode < 2; ++mode)
            for (int n = 32; n <= 256; n *= 2) {
                const uint32_t id = ((mode ? idesc_ss : idesc_ts) & ~(0x3fu << 17)) | ((uint32_t)(n >> 3) << 17);
                const long long t0 = clock64();
                for (int r = 0; r < 100; ++r)
                    for (int j = 0; j < 10; ++j) {
                        if (mode == 0)
                            mma_ts(tbase, tbase + 256 + 8 * j, make_desc(smem_u32(A_hi) + j * 1024, PANEL, 512), id, 1);
                        else
                            mma_ss(tbase, make_desc(smem_u32(A_lo) + j * 1024, PANEL, 512),
                                   make_desc(smem_u32(A_hi) + j * 1024, PANEL, 512), id, 1);
                    }
                commit(bar + 1);
                mbar_wait(bar + 1, ph);
                ph ^= 1;
                cyc[slot++] = clock64() - t0;
            }
    }
    // ---- independent accumulators: round-robin over `nd` D regions of 32 columns (N = 32)
    if (warp == 4 && lane == 0) {
        int slot = 11;
        uint32_t ph = 1;
        for (int mode = 0; mode < 2; ++mode)
            for (int nd = 1; nd <= 4; nd *= 2) {
                const uint32_t id = mode ? idesc_ss : idesc_ts;
                const long long t0 = clock64();
                for (int r = 0; r < 100; ++r)
                    for (int j = 0; j < 10; ++j)
                        for (int d = 0; d < nd; ++d) {
                            if (mode == 0)
                                mma_ts(tbase + 32 * d, tbase + 256 + 8 * j, make_desc(smem_u32(A_hi) + j * 1024, PANEL, 512), id, 1);
                            else
                                mma_ss(tbase + 32 * d, make_desc(smem_u32(A_lo) + j * 1024, PANEL, 512),
                                       make_desc(smem_u32(A_hi) + j * 1024, PANEL, 512), id, 1);
                        }
                commit(bar + 1);
                mbar_wait(bar + 1, ph);
                ph ^= 1;
                cyc[slot++] = (clock64() - t0) / nd;
            }
    }
    // ---- K-major B (SWIZZLE_128B), N = 32: TS and SS, 1 / 4 independent accumulators
    if (warp == 4 && lane == 0) {
        int slot = 17;
        uint32_t ph = 1;
        const uint32_t idk_ss = idesc_ss & ~((1u << 15) | (1u << 16));
        const uint32_t idk_ts = idk_ss;
        for (int mode = 0; mode < 2; ++mode)
            for (int nd = 1; nd <= 4; nd *= 4) {
                const long long t0 = clock64();
                for (int r = 0; r < 100; ++r)
                    for (int j = 0; j < 10; ++j) {
                        const uint32_t kb = j >> 2, ko = (j & 3) * 32;
                        for (int d = 0; d < nd; ++d) {
                            if (mode == 0)
                                mma_ts(tbase + 32 * d, tbase + 256 + 8 * j,
                                       make_desc(smem_u32(A_lo) + 3 * 16384 + kb * 4096 + ko, 16, 1024, 2), idk_ts, 1);
                            else
                                mma_ss(tbase + 32 * d, make_desc(smem_u32(A_lo) + kb * 16384 + ko, 16, 1024, 2),
                                       make_desc(smem_u32(A_lo) + 3 * 16384 + kb * 4096 + ko, 16, 1024, 2), idk_ss, 1);
                        }
                    }
                commit(bar + 1);
                mbar_wait(bar + 1, ph);
                ph ^= 1;
                cyc[slot++] = (clock64() - t0) / nd;
            }
    }
    // ---- warp-uniform issue (elect.sync), K-major B: SWIZZLE_128B (2) vs SWIZZLE_32B (6) vs none (0)
    if (warp == 4) {
        int slot = 21;
        uint32_t ph = 1;
        const uint32_t idk = idesc_ss & ~((1u << 15) | (1u << 16));
        for (int lay = 0; lay < 3; ++lay) {
            const uint32_t lt = lay == 0 ? 2u : (lay == 1 ? 6u : 0u);
            const long long t0 = clock64();
            for (int r = 0; r < 100; ++r)
                for (int j = 0; j < 10; ++j) {
                    uint64_t bd;
                    if (lay == 0) bd = make_desc(smem_u32(A_lo) + 3 * 16384 + (j >> 2) * 4096 + (j & 3) * 32, 16, 1024, lt);
                    else if (lay == 1) bd = make_desc(smem_u32(A_lo) + 3 * 16384 + j * 1024, 16, 256, lt);
                    else bd = make_desc(smem_u32(A_lo) + 3 * 16384 + j * 1024, 128, 256, lt);
                    uint32_t pred;
                    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
                    if (pred) {
                        mma_ts(tbase + 0, tbase + 256 + 8 * j, bd, idk, 1);
                        mma_ts(tbase + 32, tbase + 336 + 8 * j, bd, idk, 1);
                        mma_ts(tbase + 64, tbase + 256 + 8 * j, bd, idk, 1);
                        mma_ts(tbase + 96, tbase + 336 + 8 * j, bd, idk, 1);
                    }
                    __syncwarp();
                }
            if (lane == 0) {
                commit(bar + 1);
            }
            mbar_wait(bar + 1, ph);
            ph ^= 1;
            if (lane == 0) cyc[slot++] = (clock64() - t0) / 4;
        }
    }
    // ---- accumulator placement: four 32-column accumulators, M=128 N=32, different column offsets / issue orders
    if (warp == 4) {
        int slot = 28;
        uint32_t ph = 0;
        const uint32_t id = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t offs[4][4] = {{0, 32, 64, 96}, {0, 64, 32, 96}, {0, 64, 128, 192}, {0, 128, 32, 160}};
        for (int v = 0; v < 4; ++v) {
            const long long t0 = clock64();
            for (int r = 0; r < 100; ++r)
                for (int j = 0; j < 10; ++j) {
                    const uint64_t bd = make_desc(smem_u32(A_lo) + 3 * 16384 + j * 1024, 16, 256, 6);
                    uint32_t pred;
                    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
                    if (pred) {
                        mma_ts(tbase + offs[v][0], tbase + 256 + 8 * j, bd, id, 1);
                        mma_ts(tbase + offs[v][1], tbase + 336 + 8 * j, bd, id, 1);
                        mma_ts(tbase + offs[v][2], tbase + 256 + 8 * j, bd, id, 1);
                        mma_ts(tbase + offs[v][3], tbase + 336 + 8 * j, bd, id, 1);
                    }
                    __syncwarp();
                }
            if (lane == 0) commit(bar + 2);
            mbar_wait(bar + 2, ph);
            ph ^= 1;
            if (lane == 0) cyc[slot++] = (clock64() - t0) / 4;
        }
    }
    __syncwarp();
    // ---- cost model: uniform issue, TS, K-major SWIZZLE_32B B, 4 accumulators; M in {64,128} x N in {32,64}
    if (warp == 4) {
        int slot = 24;
        uint32_t ph = 0;
        for (int mm = 0; mm < 2; ++mm)
            for (int nn = 0; nn < 2; ++nn) {
                const uint32_t Mv = mm ? 128u : 64u, Nv = nn ? 64u : 32u;
                const uint32_t id = (1u << 4) | (2u << 7) | (2u << 10) | ((Nv >> 3) << 17) | ((Mv >> 4) << 24);
                const long long t0 = clock64();
                for (int r = 0; r < 100; ++r)
                    for (int j = 0; j < 10; ++j) {
                        const uint64_t bd = make_desc(smem_u32(A_lo) + 3 * 16384 + j * 1024, 16, 256, 6);
                        uint32_t pred;
                        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
                        if (pred) {
                            mma_ts(tbase + 0, tbase + 256 + 8 * j, bd, id, 1);
                            mma_ts(tbase + 64, tbase + 336 + 8 * j, bd, id, 1);
                            mma_ts(tbase + 128, tbase + 256 + 8 * j, bd, id, 1);
                            mma_ts(tbase + 192, tbase + 336 + 8 * j, bd, id, 1);
                        }
                        __syncwarp();
                    }
                if (lane == 0) commit(bar);
                ph ^= 1;
                mbar_wait(bar, ph);
                if (lane == 0) cyc[slot++] = (clock64() - t0) / 4;
            }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

int main()
{
    std::vector<float> mu(K * M), y(K * N);
    srand(1);
    auto rnd = [] { float s = 0; for (int i = 0; i < 12; ++i) s += rand() / (float)RAND_MAX; return s - 6.0f; };
    for (auto &v : mu) v = rnd();
    for (auto &v : y) v = rnd();
    float *dmu, *dy, *dout;
    long long *dc;
    cudaMalloc(&dmu, mu.size() * 4); cudaMalloc(&dy, y.size() * 4); cudaMalloc(&dout, 3 * M * N * 4); cudaMalloc(&dc, 256);
    cudaMemcpy(dmu, mu.data(), mu.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dy, y.data(), y.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0, 3 * M * N * 4);
    const int smem = 18 * PANEL + 256;
    cudaFuncSetAttribute(tc_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 200;
    tc_test<<<1, 192, smem>>>(dmu, dy, dout, dc, reps);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> out(3 * M * N);
    long long cyc[32];
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(cyc, dc, 256, cudaMemcpyDeviceToHost);
    printf("tmem base = 0x%llx\n", cyc[2]);
    const char *names[3] = {"3xTF32 SS", "3xTF32 TS(A_lo in TMEM)", "1xTF32"};
    for (int v = 0; v < 3; ++v) {
        double maxabs = 0, maxref = 0, fp32err = 0;
        for (int m = 0; m < (v == 1 ? 128 : M); ++m)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                float f = 0;
                for (int k = 0; k < K; ++k) {
                    ref += (double)mu[k * M + m] * (double)y[k * N + n];
                    f = fmaf(mu[k * M + m], y[k * N + n], f);
                }
                maxabs = fmax(maxabs, fabs(out[(v * M + m) * N + n] - ref));
                fp32err = fmax(fp32err, fabs((double)f - ref));
                maxref = fmax(maxref, fabs(ref));
            }
        printf("%-26s max|err| = %.3e   (fp32 FMA chain: %.3e, max|ref| = %.1f)\n", names[v], maxabs, fp32err, maxref);
    }
    for (int v = 0; v < 3; v += 1) {
        printf("variant %d, out[m][n] vs ref for m in {0,1,33,130}, n in 0..7\n", v);
        for (int m : {0, 1, 33, 130}) {
            for (int n = 0; n < 8; ++n) printf(" %9.4f", out[(v * M + m) * N + n]);
            printf("  |");
            for (int n = 0; n < 8; ++n) {
                double ref = 0;
                for (int k = 0; k < K; ++k) ref += (double)mu[k * M + m] * (double)y[k * N + n];
                printf(" %9.4f", ref);
            }
            printf("\n");
        }
    }
    const int nmma = reps * 60;
    printf("MMA stream: issue %.1f cyc/MMA, complete %.1f cyc/MMA (128x32x8 tf32; %d MMAs)\n",
           (double)cyc[0] / nmma, (double)cyc[1] / nmma, nmma);
    for (int mode = 0; mode < 2; ++mode)
        for (int i = 0; i < 4; ++i)
            printf("%s N=%3d: %.1f cyc/MMA (128xNx8)\n", mode ? "SS" : "TS", 32 << i, (double)cyc[3 + mode * 4 + i] / 1000.0);
    for (int mode = 0; mode < 2; ++mode)
        for (int i = 0; i < 3; ++i)
            printf("%s N=32, %d independent accumulators: %.1f cyc/MMA\n", mode ? "SS" : "TS", 1 << i, (double)cyc[11 + mode * 3 + i] / 1000.0);
    for (int mode = 0; mode < 2; ++mode)
        for (int i = 0; i < 2; ++i)
            printf("K-major B: %s N=32, %d independent accumulators: %.1f cyc/MMA\n", mode ? "SS" : "TS", i ? 4 : 1, (double)cyc[17 + mode * 2 + i] / 1000.0);
    const char *pl[4] = {"cols 0,32,64,96", "cols 0,64,32,96", "cols 0,64,128,192", "cols 0,128,32,160"};
    for (int i = 0; i < 4; ++i) printf("placement %s: %.1f cyc/MMA\n", pl[i], (double)cyc[28 + i] / 1000.0);
    for (int i = 0; i < 4; ++i)
        printf("cost model: TS, SWIZZLE_32B, 4 accumulators, M=%d N=%d: %.1f cyc/MMA\n", (i >> 1) ? 128 : 64, (i & 1) ? 64 : 32, (double)cyc[24 + i] / 1000.0);
    const char *ln[3] = {"SWIZZLE_128B", "SWIZZLE_32B", "no swizzle"};
    for (int i = 0; i < 3; ++i)
        printf("uniform issue, TS, K-major B %s, 4 accumulators: %.1f cyc/MMA\n", ln[i], (double)cyc[21 + i] / 1000.0);
    return 0;
}
