// mas_common.cuh -- device helpers shared by the MAS kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mas_b200.h"

// Index / protocol assertions of the debug build (-DMAS_DEBUG_CHECKS, profiles/debug_checks.sh): every
// shared-memory and output index the kernels compute is checked against its bound and a violation traps.
// compute-sanitizer is closed on the build pool (profiles/r2_sanitizer_closed.txt); this is the
// "bounds checks and asserts of your own" it asks for.  Compiled out of the product.
#ifdef MAS_DEBUG_CHECKS
#include <cstdio>
#define MAS_CHECK(cond)                                                                              \
    do {                                                                                             \
        if (!(cond)) {                                                                               \
            printf("MAS_CHECK failed: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__, \
                   (int)blockIdx.x, (int)threadIdx.x);                                               \
            __trap();                                                                                \
        }                                                                                            \
    } while (0)
#else
#define MAS_CHECK(cond) ((void)0)
#endif

namespace mas {

constexpr float kNeg = -1e9f;   // max_neg_val, core.pyx:40 (exactly representable)
constexpr int kTileY = 32;      // frames per staged tile: one 128-byte row segment per token
constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------- mbarrier (shared::cta)
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

// wait used by agents that are NOT on the critical path: back off between polls so the spin
// does not take issue slots from the warps sharing the scheduler (DP warp, MMA issuer)
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity, unsigned ns = 64)
{
    while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}

// ---------------------------------------------------------------- thread-block clusters (distributed shared memory)
// A shared::cta address is also a valid shared::cluster address of the executing CTA; map_to_cta gives the
// address of the same variable in CTA `rank` of the cluster.  The cluster-scope arrive/wait pair orders the
// remote st.shared::cluster stores issued before the arrive (same thread) before the loads after the wait.
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity, unsigned ns = 0)
{
    while (!mbar_try_wait_cluster(bar, parity))
        if (ns) __nanosleep(ns);
}
__device__ __forceinline__ void st_cluster_f4(uint32_t cluster_addr, float x, float y, float z, float w)
{
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "f"(x), "f"(y), "f"(z), "f"(w)
                 : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t cluster_addr, uint32_t v)
{
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}

__device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr));
    return v;
}

// Packed fp32 FMA (sm_100 FFMA2): (d0, d1) = (a, a) * (b0, b1) + (d0, d1), two independent
// IEEE round-to-nearest FMAs in one issue slot.  Measured on B200 (profiles/microbench/fma_pass.cu):
// the 4x8 outer-product tile reaches 74% of the fp32 peak with it, 66% with scalar FFMA.
__device__ __forceinline__ void ffma2(float &d0, float &d1, float a, float b0, float b1)
{
    asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
        "mov.b64 ra, {%2, %2};\n\t"
        "mov.b64 rb, {%3, %4};\n\t"
        "mov.b64 rc, {%0, %1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rc;\n\t}"
        : "+f"(d0), "+f"(d1)
        : "f"(a), "f"(b0), "f"(b1));
}

// ---------------------------------------------------------------- cp.async (LDGSTS)
// Asynchronous global->shared copies: no register staging, any number in flight per thread.
// `bytes` < size zero-fills the remainder (bytes == 0 reads nothing).
__device__ __forceinline__ void cp_async4(void *dst_smem, const void *src, uint32_t bytes)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst_smem)), "l"(src),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src, uint32_t bytes)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst_smem)), "l"(src),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src, uint32_t bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src),
                 "r"(bytes)
                 : "memory");
}
// The mbarrier receives one arrival once all cp.async issued so far by this thread have
// landed (.noinc: that arrival is part of the barrier's initial expected count).
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------- streaming global access
template <typename T>
__device__ __forceinline__ float load_as_f32(const T *p);

template <>
__device__ __forceinline__ float load_as_f32<float>(const float *p)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
template <>
__device__ __forceinline__ float load_as_f32<__half>(const __half *p)
{
    unsigned short u;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(u) : "l"(p));
    return __half2float(__ushort_as_half(u));
}
template <>
__device__ __forceinline__ float load_as_f32<__nv_bfloat16>(const __nv_bfloat16 *p)
{
    unsigned short u;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(u) : "l"(p));
    return __uint_as_float(static_cast<uint32_t>(u) << 16);
}
template <>
__device__ __forceinline__ float load_as_f32<double>(const double *p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return __double2float_rn(v);  // == numpy .astype(float32)
}

__device__ __forceinline__ void st_zero16(void *p)
{
    asm volatile("st.global.v4.u32 [%0], {%1, %1, %1, %1};" ::"l"(p), "r"(0u) : "memory");
}

// store the element "1" of the path dtype (esize bytes, bit pattern `one`)
__device__ __forceinline__ void st_one(void *base, int64_t elem, int esize, uint64_t one)
{
    switch (esize) {
    case 1: reinterpret_cast<uint8_t *>(base)[elem] = static_cast<uint8_t>(one); break;
    case 2: reinterpret_cast<uint16_t *>(base)[elem] = static_cast<uint16_t>(one); break;
    case 4: reinterpret_cast<uint32_t *>(base)[elem] = static_cast<uint32_t>(one); break;
    default: reinterpret_cast<uint64_t *>(base)[elem] = one; break;
    }
}

// Zero `nbytes` at `base` cooperatively: thread `tid` of `nthr`, touching only the slice
// [part, part+1)/nparts of the 16-byte body (head/tail bytes are done with part 0).
__device__ __forceinline__ void zero_fill_part(char *base, int64_t nbytes, int part, int nparts,
                                               int tid, int nthr)
{
    if (nbytes <= 0) return;
    int64_t head = (16 - (reinterpret_cast<uintptr_t>(base) & 15)) & 15;
    if (head > nbytes) head = nbytes;
    const int64_t body16 = (nbytes - head) >> 4;
    const int64_t tail0 = head + (body16 << 4);
    if (part == 0) {
        for (int64_t i = tid; i < head; i += nthr) base[i] = 0;
        for (int64_t i = tail0 + tid; i < nbytes; i += nthr) base[i] = 0;
    }
    const int64_t per = (body16 + nparts - 1) / nparts;
    int64_t lo = per * part, hi = lo + per;
    if (hi > body16) hi = body16;
    char *b16 = base + head;
    for (int64_t i = lo + tid; i < hi; i += nthr) st_zero16(b16 + (i << 4));
}

// ---------------------------------------------------------------- bulk (TMA) loads
// cp.async.bulk global -> shared::cta: the copy engine moves `bytes` (a multiple of 16, both addresses
// 16-byte aligned) and reports them to the mbarrier (complete_tx); no register staging, no LSU miss
// queue entry per 16 bytes -- a CTA with 128 staging threads keeps tens of KB in flight this way,
// which per-thread cp.async (LDGSTS) does not (measured: profiles/r2_fast3_phases.txt).
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------- TMA tensor loads
// cp.async.bulk.tensor.2d global -> shared::cta through a tensor map (CUtensorMap, built on the host by
// cuTensorMapEncodeTiled): ONE instruction moves a whole [rows x 32 frames] box, coordinates may be
// unaligned and out-of-bounds elements read as zero; completion is reported to the mbarrier as bytes.
// c0 = innermost coordinate (frame), c1 = row.
__device__ __forceinline__ void tma_load_2d(void *dst_smem, const void *tmap, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void *tmap)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// ---------------------------------------------------------------- bulk (TMA) zero fill
// cp.async.bulk shared::cta -> global: the copy engine streams a zeroed shared buffer to HBM,
// so clearing the dense output costs one instruction per `zbytes` instead of one 16-byte
// store per thread (a staging warp was spending 35% of its time on those stores).
__device__ __forceinline__ void bulk_store(void *dst_gmem, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ bool bulk_zero_ok(const void *base, int64_t nbytes)
{
    return ((reinterpret_cast<uintptr_t>(base) | static_cast<uint64_t>(nbytes)) & 15) == 0;
}
// Part `part` of `nparts` of [base, base+nbytes) (both 16-byte aligned): chunk i of the part is
// issued by thread i % nthr.  Caller commits/waits the bulk group before the region is reused.
__device__ __forceinline__ void zero_fill_bulk_part(char *base, int64_t nbytes, int part, int nparts,
                                                    const void *zbuf, int zbytes, int tid, int nthr)
{
    int64_t per = (((nbytes + nparts - 1) / nparts) + 15) & ~(int64_t)15;
    int64_t lo = per * part, hi = lo + per;
    if (hi > nbytes) hi = nbytes;
    for (int64_t o = lo + (int64_t)tid * zbytes; o < hi; o += (int64_t)nthr * zbytes)
        bulk_store(base + o, zbuf, (uint32_t)((hi - o < zbytes) ? (hi - o) : zbytes));
}

// physical float index of (token row x, frame s in tile) inside a staged tile: rows are
// 128 B, the 16-byte chunk index is XOR-ed with (x & 7) -- the TMA SWIZZLE_128B pattern, so
// the same consumer code reads tiles written by LDG/STS loaders and by cp.async.bulk.tensor.
__device__ __forceinline__ int tile_index(int x, int s)
{
    return (x << 5) + ((((s >> 2) ^ (x & 7))) << 2) + (s & 3);
}

}  // namespace mas
